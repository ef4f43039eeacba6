"""Kernel-only timing of the FVP-sum kernel(s) through the C-ABI (CUDA events around every launch, trpo_ctx_kernel_timing).

    python tools/time_fused.py [workload] [n_states] [n_fvp] [--ref]

Prints one JSON line: average kernel ms, the roofline fraction by flops_min and, with --ref, the error of
the result against the oracle on a 20 k-state prefix. Environment switches of experimental kernel variants apply."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if len(args) > 0 else "mlp64"
if "-" in name:                       # explicit shape, e.g. 17-32-32-6 (activations l t ... t l)
    layers = [int(x) for x in name.split("-")]
    ac, n_def = "l" + "t" * (len(layers) - 2) + "l", 1_000_000
else:
    layers, ac, n_def = pkg.synth.SHAPES[name]
n = int(args[1]) if len(args) > 1 else n_def
n_fvp = int(args[2]) if len(args) > 2 else 20
theta = pkg.synth.make_model(layers, 1)
rng = np.random.default_rng(2)
obs = rng.standard_normal((n, layers[0]))
std = np.exp(theta[-layers[-1]:])
v = rng.uniform(0, 1, theta.size)
flops = 6 * layers[0] * layers[1] + 10 * sum(layers[i] * layers[i + 1] for i in range(1, len(layers) - 1))
out = {"workload": name, "n": n, "env": {k: v_ for k, v_ in os.environ.items() if k.startswith("TRPO_")}}
fp32 = "--fp32" in sys.argv
with pkg.Context(layers, ac, precision=1 if fp32 else 0) as ctx:
    ctx.set_model(theta)
    ctx.set_batch(obs, std)
    for _ in range(3):
        z = ctx.fvp(v, 0.1)
    ctx.kernel_timing(True)
    for _ in range(n_fvp):
        z = ctx.fvp(v, 0.1)
    ms, k = ctx.kernel_time_ms()
    ctx.kernel_timing(False)
    out.update(kernel_ms=ms / max(k, 1), launches=k, path=ctx.path_used(),
               tflops=flops * n / (ms / max(k, 1) * 1e-3) / 1e12, frac_of_37_1=flops * n / (ms / max(k, 1) * 1e-3) / 37.1e12)
    if "--ref" in sys.argv:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import Oracle
        m = min(n, 20000)
        ctx.set_batch(obs[:m], std)
        zp = ctx.fvp(v, 0.1)
        ref = Oracle(fast=True).fvp(layers, ac, theta, std, np.ascontiguousarray(obs[:m]), 0.1, v)
        out["err_vs_oracle_prefix"] = float(np.abs(zp - ref).max() / np.abs(ref).max())
print(json.dumps(out))
