"""First-FVP cost of a Humanoid-size batch (376-256-256-17 x 1 M) staged from pinned host memory on the GEMM-chain path:
set_batch returns at once, the first FVP runs while the 3 GB cross PCIe (54 ms at 55 GB/s, tools/h2d_bandwidth.py)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from __graft_entry__ import load_package
pkg = load_package()
layers, ac = [376, 256, 256, 17], "lttl"
n = 1_000_000
theta = pkg.synth.make_model(layers, 8)
rng = np.random.default_rng(8)
obs = torch.from_numpy(rng.standard_normal((n, layers[0]))).pin_memory()
std = np.exp(theta[-layers[-1]:]); v = rng.uniform(0, 1, theta.size)
with pkg.Context(layers, ac) as ctx:
    ctx.set_model(theta)
    ctx.set_batch(obs.numpy(), std); z = ctx.fvp(v, 0.1)
    for rep in range(3):
        t0 = time.perf_counter(); ctx.set_batch(obs.numpy(), std); t1 = time.perf_counter(); z = ctx.fvp(v, 0.1); t2 = time.perf_counter()
        z2 = ctx.fvp(v, 0.1); t3 = time.perf_counter()
        print(f"set_batch {1e3*(t1-t0):.1f} ms, first fvp {1e3*(t2-t1):.1f} ms, second fvp {1e3*(t3-t2):.1f} ms, streamed vs resident max rel {np.abs(z - z2).max() / np.abs(z2).max():.2e}")
