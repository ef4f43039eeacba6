"""world_size-2 gloo test (CPU) of the multi-GPU sample sharding: each rank forms the un-normalised FVP sum of its
shard, the sums are all-reduced, and the replicated finalise / CG update reproduce the single-rank result.
The per-shard arithmetic here is the oracle's (no GPU in this container); the GPU ranks do the same through NCCL."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bench import shard_bounds
    from conftest import load_synth
    from oracle_lib import Oracle
    o = Oracle()
    s = load_synth("arm_sigma")
    L, ac, damping = s["layers"], s["acfunc"], 0.1
    N = s["Observ"].shape[0]
    lo, hi = shard_bounds(N, world, rank)
    obs = np.ascontiguousarray(s["Observ"][lo:hi])

    def fvp(v):
        # un-normalised local sum: undo the oracle's /N_local + damping*v (LogStd block handled like the GPU finalise)
        local = (o.fvp(L, ac, s["theta"], s["Std"], obs, 0.0, v)) * (hi - lo)
        t = torch.from_numpy(local)
        dist.all_reduce(t)                                  # the one collective per FVP
        z = t.numpy() / N + damping * v
        return z

    # replicated CG (TRPO_CG.c:32-107) on every rank
    b = s["b"]
    x = np.zeros_like(b); r = b.copy(); p = b.copy(); rdotr = r @ r
    for it in range(10):
        if rdotr < 1e-10:
            break
        z = fvp(p)
        v = rdotr / (p @ z)
        x += v * p; r -= v * z
        new = r @ r
        p = r + (new / rdotr) * p
        rdotr = new
    gathered = [torch.zeros(x.size, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(x))
    if rank == 0:
        np.save(out_path, np.stack([g.numpy() for g in gathered]))
    dist.destroy_process_group()


def test_two_rank_sharded_cg_matches_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_synth
    out_path = str(tmp_path / "x.npy")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    xs = np.load(out_path)
    assert np.array_equal(xs[0], xs[1])                     # replicated state stays bitwise identical across ranks
    s = load_synth("arm_sigma")
    ref = s["ref_cg"]
    assert np.abs(xs[0] - ref).max() / np.abs(ref).max() < 1e-9
