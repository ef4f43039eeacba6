"""Small driver for ncu: a few FVPs (and one CG) of a BASELINE workload through the C-ABI, nothing else.
    python tools/profile_fvp.py [workload] [n_states] [n_fvp] [cg]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "mlp64"
layers, ac, n_def = pkg.synth.SHAPES[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else n_def
n_fvp = int(sys.argv[3]) if len(sys.argv) > 3 else 3
do_cg = len(sys.argv) > 4 and sys.argv[4] == "cg"
theta = pkg.synth.make_model(layers, 1)
rng = np.random.default_rng(2)
obs = rng.standard_normal((n, layers[0]))
std = np.exp(theta[-layers[-1]:])
v = rng.uniform(0, 1, theta.size)
with pkg.Context(layers, ac) as ctx:
    ctx.set_model(theta)
    ctx.set_batch(obs, std)
    for _ in range(n_fvp):
        z = ctx.fvp(v, 0.1)
    if do_cg:
        x, info = ctx.cg(0.01 * v, 10, 0.0, 0.1)
    print("path", ctx.path_used(), "launches", ctx.launch_count(), float(np.abs(z).max()))
