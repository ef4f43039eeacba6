"""BASELINE config 5: full TRPO_Update step (policy gradient + CG + shs FVP + line search) swept over N states,
through the C-ABI with HOST buffers (set_batch + update every step). Prints one JSON line per (shape, N).
    python tools/sweep_update.py [--shapes arm,mlp64] [--sizes 10000,100000,1000000,4000000] [--cpu-sample 3000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="arm,mlp64")
ap.add_argument("--sizes", default="10000,100000,1000000,4000000")
ap.add_argument("--cpu-sample", type=int, default=3000)
args = ap.parse_args()

devnull = os.open(os.devnull, os.O_WRONLY)


def quiet(fn, *a):
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        return fn(*a)
    finally:
        os.dup2(saved, 1)


for shape in args.shapes.split(","):
    layers, ac, _ = pkg.synth.SHAPES[shape]
    theta = pkg.synth.make_model(layers, 77)
    nmax = max(int(x) for x in args.sizes.split(","))
    full = pkg.synth.make_batch(layers, ac, theta, nmax, 77)
    # CPU reference on a bounded prefix (oracle port == reference arithmetic, single thread)
    from oracle_lib import Oracle
    orc = Oracle(fast=True)
    ns = min(args.cpu_sample, nmax)
    sl = {k: (np.ascontiguousarray(v[:ns]) if k != "Std" else v) for k, v in full.items()}
    sl["Advantage"] = (sl["Advantage"] - sl["Advantage"].mean()) / sl["Advantage"].std()
    t0 = time.perf_counter()
    u_ref, info_ref = orc.update(layers, ac, theta, sl["Std"], sl["Observ"], sl["Mean"], sl["Action"], sl["Advantage"], 0.1)
    cpu_s = time.perf_counter() - t0
    with pkg.Context(layers, ac) as ctx:
        ctx.set_model(theta)
        ctx.set_batch(sl["Observ"], sl["Std"], sl["Mean"], sl["Action"], sl["Advantage"])
        u_gpu, info = quiet(ctx.update, 0.1)
        err = float(np.abs(u_gpu - u_ref).max() / np.abs(u_ref).max())
        for n in (int(x) for x in args.sizes.split(",")):
            b = {k: (np.ascontiguousarray(v[:n]) if k != "Std" else v) for k, v in full.items()}
            try:            # a training loop reuses pinned staging buffers; pageable numpy arrays halve the copy rate
                import torch
                b = {k: (torch.from_numpy(v).pin_memory().numpy() if k != "Std" else v) for k, v in b.items()}
                pinned = True
            except Exception:
                pinned = False
            times = []
            for rep in range(4):
                t0 = time.perf_counter()
                ctx.set_batch(b["Observ"], b["Std"], b["Mean"], b["Action"], b["Advantage"])
                u, info = quiet(ctx.update, 0.1)
                times.append(time.perf_counter() - t0)
            t = min(times[1:])
            print(json.dumps({"shape": shape, "layers": layers, "N": n, "update_ms_e2e": t * 1e3,
                              "states_per_s": n / t, "pinned_host_buffers": pinned, "cg_iters": info.cg_iters, "ls_steps": info.ls_steps,
                              "ls_accepted": info.ls_accepted,
                              "cpu_port_states_per_s": ns / cpu_s, "cpu_sample": ns, "cpu_cg_iters": info_ref.cg_iters,
                              "parity_on_cpu_sample_max_rel": err}), flush=True)
