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


def _loop_worker(rank, world, port, out_path):
    """Episode-sharded advantage standardisation and baseline objective: two all-reduced scalars (sum, squared
    deviations) for the advantage, one all-reduced gradient sum + one scalar for the objective -- the collectives
    trpo_vf_advantage / trpo_vf_evaluate issue on the GPU ranks."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lightweight_loop as lw
    from oracle_lib import Oracle
    o = Oracle()
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "lightweight.npz")))
    N = lw.NUM_EP * lw.EP_LEN
    ep_lo, ep_hi = rank * lw.NUM_EP // world, (rank + 1) * lw.NUM_EP // world
    lo, hi, ne = ep_lo * lw.EP_LEN, ep_hi * lw.EP_LEN, ep_hi - ep_lo
    x = np.zeros(lw.PADDED)
    x[:561] = g["x_base0"]
    obs = np.ascontiguousarray(g["it0_Observ"][lo:hi])
    base = o.vf_predict(lw.ARM_VF_LAYERS, lw.ARM_ACFUNC, x, obs, ne, lw.EP_LEN)
    # un-standardised local advantage: undo the oracle's local standardisation is not possible, so rebuild it from the
    # recurrences (same arithmetic as k_gae)
    R, V = g["it0_Reward"][lo:hi].reshape(ne, lw.EP_LEN), base.reshape(ne, lw.EP_LEN)
    ret, adv = np.zeros_like(R), np.zeros_like(R)
    acc_r, acc_a, nxt = np.zeros(ne), np.zeros(ne), np.zeros(ne)
    for t in range(lw.EP_LEN - 1, -1, -1):
        acc_r = R[:, t] + lw.GAMMA * acc_r
        acc_a = (R[:, t] + lw.GAMMA * nxt - V[:, t]) + lw.GAMMA * lw.LAM * acc_a
        ret[:, t], adv[:, t], nxt = acc_r, acc_a, V[:, t]
    ret, adv = ret.reshape(-1), adv.reshape(-1)
    s1 = torch.tensor([adv.sum()], dtype=torch.float64)
    dist.all_reduce(s1)
    mean = s1.item() / N
    s2 = torch.tensor([((adv - mean) ** 2).sum()], dtype=torch.float64)
    dist.all_reduce(s2)
    adv = (adv - mean) / np.sqrt(s2.item() / N)
    # baseline objective: local un-normalised gradient sum and squared error, all-reduced, then the replicated tail
    fx_l, g_l, pred = o.vf_evaluate(lw.ARM_VF_LAYERS, lw.ARM_ACFUNC, x, obs, ret, ne, lw.EP_LEN)
    gsum = torch.from_numpy((g_l[:561] - 0.002 * x[:561]) * (hi - lo))
    sq = torch.tensor([((pred - ret) ** 2).sum()], dtype=torch.float64)
    dist.all_reduce(gsum)
    dist.all_reduce(sq)
    grad = gsum.numpy() / N + 0.002 * x[:561]
    fx = 0.01 * sq.item() / N + 0.001 * float(x[:561] @ x[:561])
    parts = [torch.zeros(hi - lo, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(adv))
    if rank == 0:
        np.savez(out_path, adv=np.concatenate([p.numpy() for p in parts]), grad=grad, fx=np.array(fx))
    dist.destroy_process_group()


def test_two_rank_episode_sharded_advantage_and_baseline_objective(tmp_path):
    out_path = str(tmp_path / "loop.npz")
    mp.spawn(_loop_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    r = dict(np.load(out_path))
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "lightweight.npz")))
    assert np.abs(r["adv"] - g["it0_Advantage"]).max() < 1e-10
    assert np.abs(r["grad"] - g["it0_ref_evaluate_g"][:561]).max() < 1e-10 * np.abs(g["it0_ref_evaluate_g"]).max() + 1e-15
    assert abs(float(r["fx"]) - float(g["it0_ref_evaluate_fx"])) < 1e-10 * float(g["it0_ref_evaluate_fx"])
