"""10-iteration CG solve time through the device-buffer C-ABI (wall clock around N back-to-back solves + one sync).

    python tools/time_solve.py [workload] [n_states] [n_solves] [--ref]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if len(args) > 0 else "mlp64"
if "-" in name:
    layers = [int(x) for x in name.split("-")]
    ac, n_def = "l" + "t" * (len(layers) - 2) + "l", 1_000_000
else:
    layers, ac, n_def = pkg.synth.SHAPES[name]
n = int(args[1]) if len(args) > 1 else n_def
reps = int(args[2]) if len(args) > 2 else 20
theta = pkg.synth.make_model(layers, 1)
rng = np.random.default_rng(2)
obs = rng.standard_normal((n, layers[0]))
std = np.exp(theta[-layers[-1]:])
b = 0.01 * rng.standard_normal(theta.size)
dev = torch.device("cuda", 0)
d_b = torch.from_numpy(b).to(dev)
d_x = torch.zeros_like(d_b)
out = {"workload": name, "n": n, "env": {k: v for k, v in os.environ.items() if k.startswith("TRPO_")}}
with pkg.Context(layers, ac) as ctx:
    ctx.set_model(theta)
    ctx.set_batch(obs, std)
    for _ in range(3):
        ctx.cg_device(d_b.data_ptr(), d_x.data_ptr(), 10, 0.0, 0.1)
    ctx.sync()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.cg_device(d_b.data_ptr(), d_x.data_ptr(), 10, 0.0, 0.1)
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    out.update(solve_ms=dt * 1e3, launches_per_solve=(ctx.launch_count() - l0) / reps, solve_kernel=ctx.solve_kernel_used(),
               fvp_samples_per_s=10 * n / dt)
    x, info = ctx.cg(b, 10, 0.0, 0.1)
    out["cg_iters"] = info.cg_iters
    assert np.array_equal(x, d_x.cpu().numpy())
    if "--ref" in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import Oracle
        m = min(n, 20000)
        ctx.set_batch(obs[:m], std)
        xp, info = ctx.cg(b, 10, 0.0, 0.1)
        ref, nf, rd, xn = Oracle(fast=True).cg(layers, ac, theta, std, np.ascontiguousarray(obs[:m]), 0.1, b, 10, 0.0)
        out["err_vs_oracle_prefix"] = float(np.abs(xp - ref).max() / np.abs(ref).max())
        out["trace_rel"] = float(np.max(np.abs(np.array(info.cg_rdotr[:11]) - rd) / rd))
print(json.dumps(out))
