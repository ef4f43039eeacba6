#!/usr/bin/env python
"""bench.py -- headline benchmark of the natural-gradient solve (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mlp64|arm|...] [--impl reference]

A "step" is one 10-iteration conjugate-gradient solve (ResidualTh = 0, so exactly 10 Fisher-vector products,
TRPO_CG.c:45-107) over the whole rollout batch. Default workload = BASELINE.json configs[2], the configuration the
metric is quoted on: the 17-64-64-6 tanh Gaussian policy over 1M synthetic states, sharded over the N ranks
(total work fixed => "strong" scaling; one all-reduce of the P-length FVP sum per CG iteration -- by default the NVLink
peer-memory exchange inside the persistent solve kernel, `--comm nccl` for ncclAllReduce between per-iteration launches).

value    FVP samples/s with the batch already resident in HBM (10 * N_states / step time; CUDA events on the
         launching stream, L2 flushed between steps, max over ranks)
e2e      the same metric through the C-ABI with HOST buffers (trpo_ctx_set_batch + trpo_ctx_cg from pinned memory:
         H2D of the batch and b, D2H of x every step)
roofline the dominant kernel (the persistent solve kernel: 10 FVP passes per launch) timed live with CUDA events, against the
         FP64 DMMA peak measured in this process right before; traffic from the committed ncu capture (profiles/ncu_traffic.json)
parity   driver-visible correctness: sharded vs single-GPU solve (N > 1), a prefix through the unmodified reference CG() vs the GPU
cpu_baseline / --impl reference: the reference's own CPU CG (oracle/_ref, unmodified sources) on a bounded sample.
also     BASELINE configs 2, 4, 5 and the file-based drop-ins, measured in the same run (single GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CG_ITERS = 10
DAMPING = 0.1
# FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so the DMMA / DFMA pipe is measured in this process on the
# benchmarked GPU right before the timed region (trpo_probe_fp64_peak_tflops: mma.sync.m8n8k4.f64, 16 accumulators per warp).
# Fallback if the probe fails: 37.1 TFLOP/s = 148 SMs x 64 FMA/clk x 2 x 1.96 GHz (profiles/fp64_peak_r01.txt).
FP64_PEAK_FALLBACK_TFLOPS = 37.1
try:
    HBM_PEAK_GBPS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM_PEAK_GBPS = 6650.0        # fallback stated in B200_PROFILING.md


def ncu_traffic_bytes(workload, n_local, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of the same command (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep). None if
    no capture matches this (workload, states per GPU, kernel)."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    for e in table.get("entries", []):
        if e.get("workload") == workload and e.get("states_per_gpu") == n_local and e.get("kernel") == kernel:
            return e.get("dram_bytes_per_launch")
    return None


WORKLOAD_INDEX = {"arm": 1, "mlp64": 2, "pendulum64": 2, "humanoid64": 2, "humanoid256": 3}


def shard_bounds(n, world, rank):
    """Contiguous block of samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flops_min_per_sample(layers):
    """Algorithmic flops of one per-sample FVP: 6*L0*L1 + 10*sum_{i>=1} L_i*L_{i+1}."""
    return 6 * layers[0] * layers[1] + 10 * sum(layers[i] * layers[i + 1] for i in range(1, len(layers) - 1))


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in this process every 5 ms (a 10-step region of
    3 ms solves on 8 GPUs lasts ~50 ms, shorter than one `nvidia-smi -lms 100` period); `nvidia-smi` is the fallback
    when the NVML binding cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.h, self.run = index, [], None, None, None, False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def start(self):
        self.rows = []
        if self.nv:
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            try:
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(reasons(self.h))))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv:
            self.run = False
            self.t.join(timeout=2)
            sm = [r[0] for r in self.rows]
            mask = 0
            for r in self.rows:
                mask |= r[1]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(n for n, b in self.BITS.items() if mask & b), "samples": len(sm), "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


def make_workload(pkg, name, n_states):
    layers, ac, default_n = pkg.synth.SHAPES[name]
    n = n_states or default_n
    seed = pkg.synth.SEED_BASE + WORKLOAD_INDEX.get(name, 9)
    theta = pkg.synth.make_model(layers, seed)
    batch = pkg.synth.make_batch(layers, ac, theta, n, seed)
    vec = pkg.synth.make_vectors(layers, seed)
    return layers, ac, n, theta, batch, vec


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm: the unmodified reference CG() (oracle/_ref) on text files in a tmpfs directory
def reference_cg_rate(pkg, layers, ac, theta, batch, vec, n_sample, steps, warmup, tmpdir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Reference
    sl = {k: (v[:n_sample] if k != "Std" else v) for k, v in batch.items()}
    use_ref = Reference.available()
    times = []
    x_last = None
    if use_ref:
        ref = Reference(fast=True)
        mf, df = os.path.join(tmpdir, "model.txt"), os.path.join(tmpdir, "data.txt")
        pkg.textio.write_model(mf, theta)
        pkg.textio.write_data(df, sl["Mean"], sl["Std"], sl["Observ"], sl["Action"], sl["Advantage"])
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        sys.stdout.flush()
        os.dup2(devnull, 1)                      # the reference printf's a CG trace per call
        try:
            for i in range(warmup + steps):
                x_last, t = ref.cg(mf, df, layers, ac, n_sample, DAMPING, vec["b"], CG_ITERS, 0.0, 1)
                if i >= warmup:
                    times.append(t)              # the function's own returned compute seconds (file parsing excluded)
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
    else:
        orc = Oracle(fast=True)
        obs = np.ascontiguousarray(sl["Observ"])
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x_last = orc.cg(layers, ac, theta, sl["Std"], obs, DAMPING, vec["b"], CG_ITERS, 0.0)[0]
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    reference_cg_rate.last_x = np.array(x_last, copy=True) if x_last is not None else None
    return CG_ITERS * n_sample / t, t, ("reference" if use_ref else "port")


def reference_thread_scaling(pkg, layers, ac, theta, batch, vec, tmpdir, n_small=256):
    """The reference's OpenMP pragmas sit inside the per-sample layer loops (TRPO_FVP.c:790,864,889): more threads make it
    slower. One FVPFast() on a few samples at 1 thread and at all host threads documents why the arm runs single-threaded."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Reference
    if not Reference.available():
        return None
    ref = Reference(fast=True)
    sl = {k: (v[:n_small] if k != "Std" else v) for k, v in batch.items()}
    mf, df = os.path.join(tmpdir, "model_s.txt"), os.path.join(tmpdir, "data_s.txt")
    pkg.textio.write_model(mf, theta)
    pkg.textio.write_data(df, sl["Mean"], sl["Std"], sl["Observ"], sl["Action"], sl["Advantage"])
    nthreads = os.cpu_count() or 1
    out = {}
    for nt in (1, nthreads):
        _, t = ref.fvp_fast(mf, df, layers, ac, n_small, DAMPING, vec["v"], nt)
        out[nt] = n_small / t
    return {"states": n_small, "samples_per_s_1_thread": out[1], f"samples_per_s_{nthreads}_threads": out[nthreads],
            "threads": nthreads}


def run_reference_arm(args, pkg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    layers, ac, n, theta, batch, vec = make_workload(pkg, args.workload, args.states)
    # bounded sample: about 8 s of CPU work per step (the reference needs ~ 1.3 ns per flop_ref single-threaded)
    flops_ref = 10 * sum(layers[i] * layers[i + 1] for i in range(len(layers) - 1))
    n_sample = int(min(n, max(512, 8.0 / (CG_ITERS * flops_ref * 1.3e-9))))
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        rate, t_step, kind = reference_cg_rate(pkg, layers, ac, theta, batch, vec, n_sample, args.steps, max(1, min(args.warmup, 1)), tmp)
        scaling = reference_thread_scaling(pkg, layers, ac, theta, batch, vec, tmp)
    line = {
        "impl": "reference", "metric": "fvp_samples_per_sec", "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {'-'.join(map(str, layers))} policy, {n} synthetic states, "
                               f"{CG_ITERS}-iteration CG (ResidualTh=0), damping {DAMPING}"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": 1, "kind": kind,
                         "sample": f"CG() of the unmodified reference on the first {n_sample} states, NumThreads=1 "
                                   f"(its OpenMP sits inside the 64-wide layer loops and anti-scales, SURVEY.md section 6)"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if scaling:
        line["cpu_baseline"]["openmp_thread_scaling_of_FVPFast"] = scaling
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def measure_loop_body(pkg, m, steps, torch):
    """SURVEY.md section 8 rows f-3/f-4 beside the headline: everything the reference's training loop does between the
    rollouts and the next rollouts (TRPO_Lightweight.c:541-1461) on the headline batch, through the host-buffer C-ABI
    with pinned HOST arrays: stage the rollouts, baseline prediction + return + GAE + standardisation, 30 baseline
    objective/gradient evaluations (what a 25-iteration L-BFGS fit asks for; the L-BFGS vector work itself is the
    caller's libLBFGS and not in here) and the TRPO update. Wall clock, every call synchronises."""
    layers, ac, n = m["layers"], m["ac"], m["n_total"]
    ep_len = 1000
    num_ep = n // ep_len
    n = num_ep * ep_len
    b = m["batch"]
    vf_layers = [layers[0] + 1] + layers[1:-1] + [1]
    npar = sum(vf_layers[i] * vf_layers[i + 1] + vf_layers[i + 1] for i in range(len(vf_layers) - 1))
    x = np.zeros((npar + 15) // 16 * 16)
    x[:npar] = pkg.synth.make_model(vf_layers, 92)[:npar]
    rng = np.random.default_rng(5)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    obs, mean, act = pin(b["Observ"][:n]), pin(b["Mean"][:n]), pin(b["Action"][:n])
    rew = pin(-np.abs(rng.standard_normal(n)) * 3)
    devnull = os.open(os.devnull, os.O_WRONLY)
    with pkg.Context(layers, ac) as ctx, pkg.ValueFunction(ctx, vf_layers, ac) as vf:
        ctx.set_model(m["theta"])
        parts = {"stage": 0.0, "advantage": 0.0, "objective_x30": 0.0, "update": 0.0}

        def step(record):
            t0 = time.perf_counter()
            ctx.set_rollout(num_ep, ep_len, obs, b["Std"], mean, act, rew)
            ctx.sync()
            t1 = time.perf_counter()
            vf.advantage(x, n, 0.995, 0.98, fetch=False)
            t2 = time.perf_counter()
            for _ in range(30):
                vf.evaluate(x)
            t3 = time.perf_counter()
            saved = os.dup(1)                      # trpo_ctx_update prints nothing, the C host entry points may
            os.dup2(devnull, 1)
            try:
                ctx.update(DAMPING)
            finally:
                os.dup2(saved, 1)
                os.close(saved)
            t4 = time.perf_counter()
            if record:
                for k, v in zip(parts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                    parts[k] += v * 1e3 / steps
        step(False)
        l0 = ctx.launch_count()
        for _ in range(steps):
            step(True)
        launches = ctx.launch_count() - l0
    os.close(devnull)
    total = sum(parts.values())
    return {"workload": f"{'-'.join(map(str, layers))} policy + {'-'.join(map(str, vf_layers))} baseline, {n} steps "
                        f"({num_ep} episodes x {ep_len}), pinned host buffers",
            "ms_per_iteration": total, "ms": parts, "steps_per_sec": n / (total * 1e-3),
            "h2d_bytes_per_iteration": n * (layers[0] + 2 * layers[-1] + 2) * 8,
            "policy_ctx_launches_per_iteration": launches / steps}


def _quiet_fd1():
    """Context manager: send C-level stdout (the drop-ins print the reference's log lines) to /dev/null."""
    import contextlib

    @contextlib.contextmanager
    def cm():
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        sys.stdout.flush()
        os.dup2(devnull, 1)
        try:
            yield
        finally:
            os.dup2(saved, 1)
            os.close(saved)
            os.close(devnull)
    return cm()


def measure_update_sweep(pkg, torch, sizes=(10_000, 100_000, 1_000_000, 4_000_000)):
    """BASELINE configs[4]: the full TRPO update (policy gradient + CG + shs FVP + line search, TRPO_Update.c:10-1056) through
    the host-buffer C-ABI, pinned host arrays copied in every step, swept over the batch size."""
    out = {}
    for shape in ("mlp64", "arm"):
        layers, ac, _ = pkg.synth.SHAPES[shape]
        theta = pkg.synth.make_model(layers, 77)
        full = pkg.synth.make_batch(layers, ac, theta, max(sizes), 77)
        rows = []
        with pkg.Context(layers, ac) as ctx:
            ctx.set_model(theta)
            for n in sizes:
                b = {k: (torch.from_numpy(np.ascontiguousarray(v[:n])).pin_memory().numpy() if k != "Std" else v) for k, v in full.items()}
                times = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    ctx.set_batch(b["Observ"], b["Std"], b["Mean"], b["Action"], b["Advantage"])
                    with _quiet_fd1():
                        u, info = ctx.update(DAMPING)
                    times.append(time.perf_counter() - t0)
                t = min(times[1:])
                rows.append({"states": n, "update_ms_e2e": t * 1e3, "states_per_s": n / t, "cg_iters": info.cg_iters,
                             "ls_steps": info.ls_steps, "ls_accepted": info.ls_accepted})
        out[shape] = {"layers": layers, "rows": rows}
    return out


def measure_file_dropins(pkg, n=50_000):
    """e2e of the FILE-based drop-in CG_GPU (the FVP_FPGA / CG_FPGA replacement, TRPO.h:98,101): wall clock of the whole call
    (text parse or binary read, pinned staging, solve, release) and the compute seconds the call returns, for the reference's
    text DataFile and for the binary batch file; plus the host-buffer API from PAGEABLE memory."""
    layers, ac, _ = pkg.synth.SHAPES["mlp64"]
    theta = pkg.synth.make_model(layers, 5)
    batch = pkg.synth.make_batch(layers, ac, theta, n, 5)
    vec = pkg.synth.make_vectors(layers, 5)
    res = {"workload": f"mlp64: 17-64-64-6 policy, {n} states, 10-iteration CG (ResidualTh=0)"}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        mf, tf, bf = os.path.join(tmp, "m.txt"), os.path.join(tmp, "d.txt"), os.path.join(tmp, "d.bin")
        pkg.textio.write_model(mf, theta)
        pkg.textio.write_data(tf, batch["Mean"], batch["Std"], batch["Observ"], batch["Action"], batch["Advantage"])
        pkg.api.batch_file_write(bf, batch["Observ"], batch["Std"], batch["Mean"], batch["Action"], batch["Advantage"])
        xs = {}
        for tag, df in (("text_datafile", tf), ("binary_datafile", bf)):
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                with _quiet_fd1():
                    x, t_ret = pkg.CG_GPU(mf, df, layers, ac, n, DAMPING, vec["b"], CG_ITERS, 0.0, 1)
                wall = time.perf_counter() - t0
                if best is None or wall < best[0]:
                    best = (wall, t_ret)
            xs[tag] = x
            res[tag] = {"wall_ms_whole_call": best[0] * 1e3, "returned_compute_ms": best[1] * 1e3,
                        "fvp_samples_per_s_whole_call": CG_ITERS * n / best[0], "file_bytes": os.path.getsize(df)}
        res["text_vs_binary_bitwise_equal"] = bool(np.array_equal(xs["text_datafile"], xs["binary_datafile"]))
    with pkg.Context(layers, ac) as ctx:
        ctx.set_model(theta)
        obs = np.ascontiguousarray(batch["Observ"])              # plain numpy = pageable host memory
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            ctx.set_batch(obs, batch["Std"])
            ctx.cg(vec["b"], CG_ITERS, 0.0, DAMPING)
            ts.append(time.perf_counter() - t0)
        res["host_api_pageable_buffers"] = {"ms_per_step": min(ts[1:]) * 1e3, "fvp_samples_per_s": CG_ITERS * n / min(ts[1:])}
    return res


def run_gpu_arm(args, pkg):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = pkg.api.lib()
    # roofline denominators measured here and now, on this GPU
    fp64_peak = L.trpo_probe_fp64_peak_tflops(local_rank)
    fp64_src = "measured in this run: mma.sync.m8n8k4.f64 (DMMA, the pipe DFMA shares), 16 accumulators/warp, trpo_probe_fp64_peak_tflops"
    if not fp64_peak or fp64_peak <= 0:
        fp64_peak, fp64_src = FP64_PEAK_FALLBACK_TFLOPS, "fallback: profiles/fp64_peak_r01.txt (the in-run probe failed)"
    tf32_peak = L.trpo_probe_tf32_mma_sync_tflops(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def attach_comm(ctx):
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(pkg.api.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.init_comm(bytes(uid.cpu().numpy().tobytes()), rank, world)
        if args.comm == "p2p":
            # peer-memory all-reduce fused into our kernels: exchange the CUDA IPC handles of the comm buffers
            mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).to(dev)
            allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(allh, mine)
            ctx.p2p_attach(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
            dist.barrier()

    def measure(workload, states, steps, warmup, with_clocks, precision="fp64", keep_ctx=False):
        """One workload: resident leg (value), end-to-end leg (e2e) and the dominant-kernel roofline. Returns a dict."""
        layers, ac, n_total, theta, batch, vec = make_workload(pkg, workload, states)
        lo, hi = shard_bounds(n_total, world, rank)
        n_local = hi - lo
        P = len(theta)
        # pinned host staging buffers (the e2e leg copies from these every step)
        obs_pin = torch.from_numpy(np.ascontiguousarray(batch["Observ"][lo:hi])).pin_memory()
        b_pin = torch.from_numpy(vec["b"].copy()).pin_memory()
        x_pin = torch.zeros(P, dtype=torch.float64).pin_memory()
        stream = torch.cuda.Stream(device=dev)
        ctx = pkg.Context(layers, ac, device=local_rank, precision=1 if precision == "fp32" else 0)
        ctx.set_stream(stream.cuda_stream)
        if args.path:
            ctx.set_path({"chain": pkg.api.PATH_GEMM_CHAIN, "fused": pkg.api.PATH_FUSED}[args.path])
        if world > 1:
            attach_comm(ctx)
        ctx.set_model(theta)
        ctx.set_batch(obs_pin.numpy(), batch["Std"])
        assert ctx.global_samples() == n_total
        d_b = torch.from_numpy(vec["b"]).to(dev)
        d_x = torch.zeros(P, dtype=torch.float64, device=dev)

        def timed_loop(step_fn, nsteps, nwarm):
            with torch.cuda.stream(stream):
                for _ in range(nwarm):
                    flush_buf.zero_()
                    step_fn()
                barrier()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
                l0 = ctx.launch_count()
                t_wall = time.perf_counter()
                for e0, e1 in evs:
                    flush_buf.zero_()              # L2 flush between timed steps (outside the event pair)
                    e0.record(stream)
                    step_fn()
                    e1.record(stream)
                barrier()
                t_wall = time.perf_counter() - t_wall
                l1 = ctx.launch_count()
            ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) / nsteps, (l1 - l0), t_wall

        # ---- leg 1: device-resident (value) ---------------------------------------------------------------------
        def step_resident():
            ctx.cg_device(d_b.data_ptr(), d_x.data_ptr(), CG_ITERS, 0.0, DAMPING)

        ctx.kernel_timing(True)
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step_resident()
            torch.cuda.synchronize()
        ctx.kernel_time_ms()
        sampler = ClockSampler(local_rank)
        if with_clocks:
            sampler.start()
        ms_step, launches, _ = timed_loop(step_resident, steps, 0)
        clocks = sampler.stop() if with_clocks else None
        k_ms, k_n = ctx.kernel_time_ms()
        ctx.kernel_timing(False)
        solve_kernel = ctx.solve_kernel_used()
        # the same region without the per-kernel event pairs (the per-iteration launch path replays a CUDA graph then)
        ms_step_graph, _, _ = timed_loop(step_resident, steps, 2)
        x_resident = d_x.cpu().numpy().copy()

        # ---- leg 2: end to end through the host-buffer C-ABI (e2e) -----------------------------------------------
        import ctypes as C

        def step_e2e():
            ctx.set_batch(obs_pin.numpy(), batch["Std"])                                  # H2D of the rollout batch
            rc = L.trpo_ctx_cg(ctx.h, C.cast(b_pin.data_ptr(), pkg.api.c_double_p),
                               C.cast(x_pin.data_ptr(), pkg.api.c_double_p), CG_ITERS, 0.0, DAMPING)
            if rc:
                raise RuntimeError(pkg.api.last_error())

        ms_e2e, _, wall_e2e = timed_loop(step_e2e, steps, warmup)
        # the host-buffer calls synchronise internally, so wall clock per step is the honest end-to-end figure
        e2e_ms = max(ms_e2e, wall_e2e * 1e3 / steps)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        A = layers[-1]
        path_used = {1: "gemm_chain", 2: "fused_dmma"}.get(ctx.path_used(), "?")
        assert np.isfinite(x_resident).all() and np.isfinite(x_pin.numpy()).all()
        assert ctx.comm_error() == 0, "a peer-memory wait timed out"
        if path_used == "fused_dmma":
            assert np.array_equal(x_resident, x_pin.numpy()), "resident and host-buffer solves differ"
        else:   # GEMM chain: the first FVP of a streamed batch walks it piece by piece (same sums, grouped differently)
            assert np.abs(x_resident - x_pin.numpy()).max() <= 1e-10 * np.abs(x_resident).max(), "resident and host-buffer solves differ"
        fl = flops_min_per_sample(layers)
        # dominant kernel: the persistent solve kernel runs all CG_ITERS FVP passes in one launch, otherwise one FVP per launch
        fvps_per_launch = CG_ITERS if solve_kernel else 1
        kernel_name = "k_cg_solve (persistent: %d FVP passes + reductions + CG update per launch)" % CG_ITERS if solve_kernel \
            else ("k_fvp_fused / k_fvp_warp" if path_used == "fused_dmma" else "GEMM chain (TMA-fed k_fwd_tma / k_bwd_tma / k_outer_tma + k_chain_tail), all launches of one FVP")
        k_avg_ms = k_ms / max(k_n, 1)
        achieved = fvps_per_launch * fl * n_local / (k_avg_ms * 1e-3) / 1e12 if k_n else None
        if precision == "fp32":
            peak, peak_src, bound_note = (tf32_peak / 3.0 if tf32_peak and tf32_peak > 0 else None,
                                          "measured in this run: mma.sync.m16n8k8 TF32 throughput / 3 (a 3xTF32 product is three tensor-core products)",
                                          "FP32 mode: tensor pipe, 3xTF32")
        else:
            peak, peak_src, bound_note = fp64_peak, fp64_src, "FP64 pipe (DMMA)"
        out = {
            "layers": layers, "n_total": n_total, "n_local": n_local, "P": P, "ms_step": ms_step, "ms_step_graph": ms_step_graph,
            "launches": int(launches), "clocks": clocks, "solve_kernel": solve_kernel,
            "value": CG_ITERS * n_total / (ms_step * 1e-3), "e2e_ms": e2e_ms,
            "e2e_value": CG_ITERS * n_total / (e2e_ms * 1e-3),
            "h2d": n_local * layers[0] * 8 + 2 * A * 8 + P * 8, "d2h": P * 8 + 576, "path": path_used,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if (achieved and peak) else None,
                         "traffic": ncu_traffic_bytes(workload, n_local, "k_cg_solve" if solve_kernel else path_used),
                         "kernel": kernel_name, "pipe": bound_note, "kernel_avg_ms": k_avg_ms, "kernel_launches_timed": k_n,
                         "fvp_passes_per_launch": fvps_per_launch, "flops_per_sample": fl,
                         "hbm_context": {"algorithmic_bytes_per_launch": fvps_per_launch * n_local * layers[0] * 8,
                                         "achieved_GBps": (fvps_per_launch * n_local * layers[0] * 8 / (k_avg_ms * 1e-3) / 1e9) if k_n else None,
                                         "measured_peak_GBps": HBM_PEAK_GBPS,
                                         "note": f"compute bound by design: {fl / (8 * layers[0]):.0f} flop/B against a ridge of 5.7"},
                         "peak_source": peak_src},
            "theta": theta, "batch": batch, "vec": vec, "ac": ac, "x": x_resident,
        }
        if keep_ctx:
            out["ctx"] = ctx
        else:
            ctx.close()
        return out

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    m = measure(args.workload, args.states, args.steps, args.warmup, True, precision=args.precision)
    layers, n_total, P = m["layers"], m["n_total"], m["P"]

    # ---- parity, visible to the driver --------------------------------------------------------------------------------
    parity = {}
    cpu_line = None
    if rank == 0:
        theta, batch, vec, ac = m["theta"], m["batch"], m["vec"], m["ac"]
        if world > 1:
            # the sharded solve against the same solve on ONE GPU over the whole batch (different reduction tree: ~1e-13)
            with pkg.Context(layers, ac, device=local_rank, precision=1 if args.precision == "fp32" else 0) as c1:
                if args.path:
                    c1.set_path({"chain": pkg.api.PATH_GEMM_CHAIN, "fused": pkg.api.PATH_FUSED}[args.path])
                c1.set_model(theta)
                c1.set_batch(batch["Observ"], batch["Std"])
                x1, _ = c1.cg(vec["b"], CG_ITERS, 0.0, DAMPING)
            parity["vs_single_gpu_max_rel"] = float(np.abs(m["x"] - x1).max() / np.abs(x1).max())
        if not args.no_cpu_baseline:
            # a prefix of the batch through the UNMODIFIED reference CG() on the host and through the GPU path
            flops_ref = 10 * sum(layers[i] * layers[i + 1] for i in range(len(layers) - 1))
            n_sample = int(min(n_total, max(512, 12.0 / (CG_ITERS * flops_ref * 1.3e-9))))
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
                rate, t_step, kind = reference_cg_rate(pkg, layers, ac, theta, batch, vec, n_sample, 1, 0, tmp)
            cpu_line = {"value": rate, "unit": "samples/s", "cores": 1, "kind": kind,
                        "sample": f"one {CG_ITERS}-iteration CG() of the unmodified reference on the first {n_sample} "
                                  f"states ({t_step:.1f} s), NumThreads=1 of {os.cpu_count()} host cores"}
            x_ref = reference_cg_rate.last_x
            with pkg.Context(layers, ac, device=local_rank, precision=1 if args.precision == "fp32" else 0) as c1:
                c1.set_model(theta)
                c1.set_batch(np.ascontiguousarray(batch["Observ"][:n_sample]), batch["Std"])
                xg, _ = c1.cg(vec["b"], CG_ITERS, 0.0, DAMPING)
            parity["prefix_vs_reference_max_rel"] = float(np.abs(xg - x_ref).max() / np.abs(x_ref).max())
            parity["prefix_states"] = n_sample
            parity["tolerance"] = "FP64 build: FVP 1e-10, CG 1e-8 (tests/test_gpu_parity.py)" if args.precision == "fp64" else "FP32 mode: FVP 1e-4"
    barrier()

    also = {}
    if args.workload == "mlp64" and not args.states and not args.no_secondary and args.precision == "fp64":
        # BASELINE configs[1] (armDOF_0 policy, 50 k states) measured in the same run, reported beside the headline
        a = measure("arm", 0, args.steps, args.warmup, False)
        also["arm_50k"] = {"workload": "arm: 15-16-16-3 policy, 50000 synthetic states, 10-iteration CG",
                           "value": CG_ITERS * a["n_total"] / (min(a["ms_step_graph"], a["ms_step"]) * 1e-3), "unit": "samples/s",
                           "cg_solve_ms": min(a["ms_step_graph"], a["ms_step"]), "cg_solve_ms_with_kernel_event_pairs": a["ms_step"],
                           "e2e_value": a["e2e_value"], "e2e_ms_per_step": a["e2e_ms"], "kernel_path": a["path"],
                           "persistent_solve_kernel": a["solve_kernel"],
                           "roofline_frac": a["roofline"]["frac"], "kernel_avg_ms": a["roofline"]["kernel_avg_ms"]}
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            # the reference's CPU CG() on the whole 50 k-state arm batch (about 4 s on one core)
            with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
                rate, t_step, kind = reference_cg_rate(pkg, a["layers"], a["ac"], a["theta"], a["batch"], a["vec"],
                                                       a["n_total"], 1, 0, tmp)
            also["arm_50k"]["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": 1, "kind": kind,
                                               "sample": f"one 10-iteration CG() of the unmodified reference on all "
                                                         f"{a['n_total']} states ({t_step:.1f} s), NumThreads=1"}
            also["arm_50k"]["parity_vs_reference_max_rel"] = float(np.abs(a["x"] - reference_cg_rate.last_x).max() / np.abs(reference_cg_rate.last_x).max())
        if world == 1:
            # the arm policy when the kernel is fed: 1 M states
            a1 = measure("arm", 1_000_000, max(2, args.steps // 2), 3, False)
            also["arm_1m"] = {"workload": "arm: 15-16-16-3 policy, 1000000 synthetic states, 10-iteration CG",
                              "value": a1["value"], "unit": "samples/s", "cg_solve_ms": a1["ms_step"], "e2e_value": a1["e2e_value"],
                              "roofline_frac": a1["roofline"]["frac"], "kernel_avg_ms": a1["roofline"]["kernel_avg_ms"],
                              "flops_per_sample": a1["roofline"]["flops_per_sample"]}
            also["loop_body_1m"] = measure_loop_body(pkg, m, args.steps, torch)
            # BASELINE configs[3]: Humanoid-size policy (376-256-256-17), 1 M states, FP64 and the FP32 mode side by side
            st_h = max(2, args.steps // 5)
            h64 = measure("humanoid256", 1_000_000, st_h, 3, False, keep_ctx=True)
            h32 = measure("humanoid256", 1_000_000, st_h, 3, False, precision="fp32", keep_ctx=True)
            z64 = h64["ctx"].fvp(h64["vec"]["v"], DAMPING)
            z32 = h32["ctx"].fvp(h32["vec"]["v"], DAMPING)
            h64["ctx"].close(); h32["ctx"].close()
            err = np.abs(z32 - z64)
            also["humanoid256_1m"] = {
                "workload": "humanoid256: 376-256-256-17 policy, 1000000 synthetic states, 10-iteration CG (GEMM-chain path)",
                "fp64": {"value": h64["value"], "unit": "samples/s", "cg_solve_ms": h64["ms_step"], "fvp_ms": h64["roofline"]["kernel_avg_ms"],
                         "e2e_value": h64["e2e_value"], "e2e_ms_per_step": h64["e2e_ms"], "roofline_frac": h64["roofline"]["frac"],
                         "roofline_peak_tflops": h64["roofline"]["peak"],
                         # dram__bytes of ONE FVP (all kernels of the chain; profiles/ncu_traffic.json) against the algorithmic 8*L0*N
                         "dram_bytes_per_fvp": h64["roofline"]["traffic"],
                         "dram_bytes_over_algorithmic": (h64["roofline"]["traffic"] / (8.0 * 376 * h64["n_local"])) if h64["roofline"]["traffic"] else None},
                "fp32_mode": {"value": h32["value"], "unit": "samples/s", "cg_solve_ms": h32["ms_step"], "fvp_ms": h32["roofline"]["kernel_avg_ms"],
                              "e2e_value": h32["e2e_value"], "roofline_frac_of_3xTF32": h32["roofline"]["frac"],
                              "roofline_peak_tflops": h32["roofline"]["peak"], "kernel": h32["roofline"]["pipe"]},
                "fp32_vs_fp64_speedup_fvp": h64["roofline"]["kernel_avg_ms"] / h32["roofline"]["kernel_avg_ms"],
                "fp32_fvp_error_vs_fp64_gpu_1m_states": {"max_abs_over_max_ref": float(err.max() / np.abs(z64).max()),
                                                         "rel_l2": float(np.linalg.norm(err) / np.linalg.norm(z64)),
                                                         "stated_tolerance": 1e-4},
                "flops_per_sample": h64["roofline"]["flops_per_sample"]}
            del h64, h32
            also["update_sweep"] = measure_update_sweep(pkg, torch)
            also["file_dropins_50k"] = measure_file_dropins(pkg)

    if rank == 0:
        line = {
            "metric": "fvp_samples_per_sec", "value": m["value"], "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_step"], "cg_solve_ms": m["ms_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.precision == "fp64" else "f32 (3xTF32 products, f32 slice sums, f64 reduction + CG)",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {'-'.join(map(str, layers))} policy, {n_total} synthetic states, "
                                   f"{CG_ITERS}-iteration CG (ResidualTh=0), damping {DAMPING}",
                       "kernel_path": m["path"] + (" / persistent cooperative solve kernel" if m["solve_kernel"] else ""),
                       "l2": "flushed between steps (256 MiB write); the 1M-state batch (136 MB) exceeds L2",
                       "parallelism": f"samples sharded over {world} GPU(s), 1 all-reduce of P={P} doubles per FVP"
                                      + (f" ({args.comm}" + (": peer-memory exchange inside the solve kernel)" if m["solve_kernel"] else ")") if world > 1 else "")},
            "clocks": m["clocks"],
            "e2e": {"value": m["e2e_value"], "unit": "samples/s", "ms_per_step": m["e2e_ms"],
                    "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
            "gpu_launches": m["launches"], "cg_solve_ms_without_kernel_event_pairs": m["ms_step_graph"],
            "roofline": m["roofline"],
            "parity": parity,
        }
        if also:
            line["also"] = also
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mlp64")
    ap.add_argument("--states", type=int, default=0, help="override the number of synthetic states")
    ap.add_argument("--path", default="", choices=["", "chain", "fused"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp32 = the optional 3xTF32 mode of the FVP (stated tolerance 1e-4); the headline is fp64")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the arm-50k secondary measurement")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU FVP-sum all-reduce: NVLink peer-memory exchange inside our kernels (default: lets the whole "
                         "solve run as one persistent kernel per GPU) or ncclAllReduce between per-iteration launches")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    from __graft_entry__ import load_package
    pkg = load_package()
    # Only the JSON line may reach stdout: NCCL / the reference's printf write to fd 1, so point fd 1 at stderr for the
    # run and print the result through a private copy of the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference_arm(args, pkg)
    else:
        run_gpu_arm(args, pkg)
    real_stdout.flush()


if __name__ == "__main__":
    main()
