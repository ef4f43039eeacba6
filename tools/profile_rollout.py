"""Rows f-3 / f-4 at BASELINE sizes: one whole training-loop iteration after the rollouts, through the C-ABI with pinned
HOST buffers -- stage the rollouts, baseline prediction + return + GAE + standardisation, baseline objective/gradient
(the libLBFGS callback), a full 25-iteration baseline fit by the reference's vendored libLBFGS when oracle/_ref is there,
and the TRPO update. Prints one JSON line per (shape, N) with the CPU reference timed beside it on a bounded sample.
    python tools/profile_rollout.py [--shapes arm,mlp64] [--sizes 3000,50000,1000000] [--cpu-sample 3000]"""
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
from oracle_lib import Oracle, Reference  # noqa: E402

pkg = load_package()
ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="arm,mlp64")
ap.add_argument("--sizes", default="3000,50000,1000000")
ap.add_argument("--cpu-sample", type=int, default=3000)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
GAMMA, LAM = 0.995, 0.98

devnull = os.open(os.devnull, os.O_WRONLY)


def quiet(fn, *a, **k):
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        return fn(*a, **k)
    finally:
        os.dup2(saved, 1)
        os.close(saved)


def pin(a):
    try:
        import torch
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    except Exception:
        return np.ascontiguousarray(a)


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


orc = Oracle(fast=True)
ref = Reference(fast=True) if Reference.available() else None
for shape in args.shapes.split(","):
    layers, ac, _ = pkg.synth.SHAPES[shape]
    O, A = layers[0], layers[-1]
    vf_layers = [O + 1] + layers[1:-1] + [1]
    npar = sum(vf_layers[i] * vf_layers[i + 1] + vf_layers[i + 1] for i in range(len(vf_layers) - 1))
    padded = (npar + 15) // 16 * 16
    theta = pkg.synth.make_model(layers, 91)
    x0 = np.zeros(padded)
    x0[:npar] = pkg.synth.make_model(vf_layers, 92)[:npar]
    for n in (int(s) for s in args.sizes.split(",")):
        ep_len = 150 if n % 150 == 0 else 1000
        num_ep = n // ep_len
        n = num_ep * ep_len
        rng = np.random.default_rng(n)
        b = pkg.synth.make_batch(layers, ac, theta, n, 91)
        reward = -np.abs(rng.normal(size=n)) * 3
        hb = {k: (pin(v) if k != "Std" else v) for k, v in b.items()}
        h_reward = pin(reward)
        out = dict(shape=shape, layers=layers, vf_layers=vf_layers, N=n, ep_len=ep_len)
        with pkg.Context(layers, ac) as ctx, pkg.ValueFunction(ctx, vf_layers, ac) as vf:
            ctx.set_model(theta)
            stage = lambda: (ctx.set_rollout(num_ep, ep_len, hb["Observ"], hb["Std"], hb["Mean"], hb["Action"], h_reward), ctx.sync())
            out["stage_ms"] = timed(stage, args.reps)
            if layers[0] == 15 and layers[-1] == 3:
                # the device-side producer instead of staging host rollouts: counter-based generator / host rand() draws
                out["rollout_arm_device_rng_ms"] = timed(lambda: (ctx.rollout_arm(num_ep, ep_len, None, 1), ctx.sync()), args.reps)
                draws = np.random.default_rng(3).integers(0, 2**31 - 1, num_ep * (3 + 6 * ep_len), dtype=np.int32)
                out["rollout_arm_host_draws_ms"] = timed(lambda: (ctx.rollout_arm(num_ep, ep_len, draws, 0), ctx.sync()), args.reps)
                stage()
            out["advantage_ms"] = timed(lambda: vf.advantage(x0, n, GAMMA, LAM, fetch=False), args.reps)
            out["vf_evaluate_ms"] = timed(lambda: vf.evaluate(x0), args.reps)
            if ref is not None:
                t0 = time.perf_counter()
                xf, fx, rc = ref.lbfgs(x0, vf.callback_pointer(), max_iterations=25)
                out["vf_fit_lbfgs25_ms"] = (time.perf_counter() - t0) * 1e3
                out["vf_fit_fx"] = [vf.evaluate(x0)[0], fx]
            out["update_ms"] = timed(lambda: quiet(ctx.update, 0.1), args.reps)

            def iteration():
                ctx.set_rollout(num_ep, ep_len, hb["Observ"], hb["Std"], hb["Mean"], hb["Action"], h_reward)
                vf.advantage(x0, n, GAMMA, LAM, fetch=False)
                if ref is not None:
                    ref.lbfgs(x0, vf.callback_pointer(), max_iterations=25)
                quiet(ctx.update, 0.1)
            out["iteration_ms"] = timed(iteration, max(1, args.reps // 2))
            # parity on a bounded prefix + the CPU reference timed on it (single thread, the reference's arithmetic)
            ns = min(args.cpu_sample, n) // ep_len * ep_len or ep_len
            ne = ns // ep_len
            sl = {k: (np.ascontiguousarray(v[:ns]) if k != "Std" else v) for k, v in b.items()}
            ctx.set_rollout(ne, ep_len, sl["Observ"], sl["Std"], sl["Mean"], sl["Action"], reward[:ns])
            ret, adv = vf.advantage(x0, ns, GAMMA, LAM)
            fx, g = vf.evaluate(x0)
            t0 = time.perf_counter()
            base = orc.vf_predict(vf_layers, ac, x0, sl["Observ"], ne, ep_len)
            r_ref, a_ref = orc.gae(reward[:ns], base, ne, ep_len, GAMMA, LAM)
            cpu_adv = time.perf_counter() - t0
            t0 = time.perf_counter()
            if ref is not None:
                f_ref, g_ref, _ = ref.vf_evaluate(vf_layers, ac, x0, sl["Observ"], r_ref, ne, ep_len)
                kind = "reference"
            else:
                f_ref, g_ref, _ = orc.vf_evaluate(vf_layers, ac, x0, sl["Observ"], r_ref, ne, ep_len)
                kind = "port"
            cpu_eval = time.perf_counter() - t0
            out["parity"] = dict(sample=ns, ret=float(np.abs(ret - r_ref).max() / np.abs(r_ref).max()),
                                 adv=float(np.abs(adv - a_ref).max() / np.abs(a_ref).max()),
                                 fx=float(abs(fx - f_ref) / abs(f_ref)), g=float(np.abs(g - g_ref).max() / np.abs(g_ref).max()))
            out["cpu"] = dict(kind=kind, cores=1, sample=ns, advantage_ms_scaled=cpu_adv * 1e3 * n / ns,
                              vf_evaluate_ms_scaled=cpu_eval * 1e3 * n / ns)
            out["launches"] = int(ctx.launch_count())
        print(json.dumps(out), flush=True)
