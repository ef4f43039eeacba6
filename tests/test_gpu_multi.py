"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; the 4- and
8-rank cases run when the box has that many): samples sharded over the ranks, FVP sums all-reduced by ncclAllReduce and
by the fused peer-memory kernels; both must reproduce the single-batch reference result and leave bitwise identical CG
state on every rank. bench.py repeats the check at full size on every multi-GPU run (`parity` in its JSON line)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, name):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from __graft_entry__ import load_package
    from bench import shard_bounds
    from conftest import load_synth
    pkg = load_package()
    s = load_synth(name)
    L, ac = s["layers"], s["acfunc"]
    N = s["Observ"].shape[0]
    lo, hi = shard_bounds(N, world, rank)
    ctx = pkg.Context(L, ac, device=rank)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pkg.api.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.init_comm(bytes(uid.cpu().numpy().tobytes()), rank, world)
    ctx.set_model(s["theta"])
    ctx.set_batch(s["Observ"][lo:hi], s["Std"], s["Mean"][lo:hi], s["Action"][lo:hi], s["Advantage"][lo:hi])
    assert ctx.global_samples() == N
    out = {}
    out["fvp_nccl"] = ctx.fvp(s["v"], 0.1)
    out["cg_nccl"], info = ctx.cg(s["b"], 10, 1e-10, 0.1)
    out["iters_nccl"] = np.array(info.cg_iters)
    out["upd_nccl"], _ = ctx.update(0.1)
    mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).to(dev)
    allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    ctx.p2p_attach(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    dist.barrier()
    out["fvp_p2p"] = ctx.fvp(s["v"], 0.1)
    out["cg_p2p"], info = ctx.cg(s["b"], 10, 1e-10, 0.1)
    out["iters_p2p"] = np.array(info.cg_iters)
    out["fvp_p2p_again"] = ctx.fvp(s["v"], 0.1)
    out["upd_p2p"], _ = ctx.update(0.1)
    out["comm_error"] = np.array(ctx.comm_error())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("name", ["arm_sigma", "mlp64", "acts5"])
def test_sharded_solve(tmp_path, name, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from conftest import load_synth, rel_err
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), name), nprocs=world, join=True)
    ranks = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]
    r0 = ranks[0]
    s = load_synth(name)
    for r in ranks[1:]:
        for k in r0:
            assert np.array_equal(r0[k], r[k]), f"{k} differs between ranks"
    assert r0["comm_error"] == 0
    for tag in ("nccl", "p2p"):
        assert rel_err(r0[f"fvp_{tag}"], s["ref_fvpfast"])[0] < 1e-10
        assert rel_err(r0[f"cg_{tag}"], s["ref_cg"])[0] < 1e-8
        assert rel_err(r0[f"upd_{tag}"], s["ref_update"])[0] < 1e-8
    assert np.array_equal(r0["fvp_p2p"], r0["fvp_p2p_again"])
    if world == 2:
        assert np.array_equal(r0["fvp_p2p"], r0["fvp_nccl"])   # 2 ranks: a+b in either order is the same double
    else:
        assert rel_err(r0["fvp_p2p"], r0["fvp_nccl"])[0] < 1e-13   # different reduction trees


def _loop_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import lightweight_loop as lw
    from __graft_entry__ import load_package
    pkg = load_package()
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "lightweight.npz")))
    ep_lo, ep_hi = rank * lw.NUM_EP // world, (rank + 1) * lw.NUM_EP // world      # whole episodes per rank
    lo, hi = ep_lo * lw.EP_LEN, ep_hi * lw.EP_LEN
    ctx = pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC, device=rank)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pkg.api.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.init_comm(bytes(uid.cpu().numpy().tobytes()), rank, world)
    vf = pkg.ValueFunction(ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC)
    x = np.zeros(lw.PADDED)
    x[:561] = g["x_base0"]
    out = {}
    ctx.set_model(g["theta0"])
    ctx.set_rollout(ep_hi - ep_lo, lw.EP_LEN, g["it0_Observ"][lo:hi], g["it0_Std"], g["it0_Mean"][lo:hi],
                    g["it0_Action"][lo:hi], g["it0_Reward"][lo:hi])
    out["ret"], out["adv"] = vf.advantage(x, hi - lo, lw.GAMMA, lw.LAM)
    fx, out["g"] = vf.evaluate(x)
    out["fx"] = np.array(fx)
    out["theta1"], _ = ctx.update(0.1)
    # the policy's FVP sums over peer memory while the baseline keeps NCCL on the same stream
    mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).to(dev)
    allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    ctx.p2p_attach(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    dist.barrier()
    fx2, g2 = vf.evaluate(x)
    out["fx_again"], out["g_again"] = np.array(fx2), g2
    out["theta1_p2p"], _ = ctx.update(0.1)
    out["comm_error"] = np.array(ctx.comm_error())
    np.savez(os.path.join(out_dir, f"loop{rank}.npz"), **out)
    dist.barrier()
    vf.close()
    ctx.close()
    dist.destroy_process_group()


def test_two_gpu_sharded_training_loop_body(tmp_path):
    """Episodes sharded over 2 ranks: advantage standardised over the global batch, baseline objective / gradient and the
    TRPO update all-reduced; every rank gets the single-GPU (= reference) values."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from conftest import rel_err
    mp.spawn(_loop_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (dict(np.load(tmp_path / f"loop{r}.npz")) for r in range(2))
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "lightweight.npz")))
    for k in ("fx", "g", "theta1", "theta1_p2p", "fx_again", "g_again"):
        assert np.array_equal(r0[k], r1[k]), f"{k} differs between ranks"
    assert r0["comm_error"] == 0 and r1["comm_error"] == 0
    assert rel_err(np.concatenate([r0["ret"], r1["ret"]]), g["it0_Return"])[0] < 1e-10
    assert rel_err(np.concatenate([r0["adv"], r1["adv"]]), g["it0_Advantage"])[0] < 1e-10
    assert abs(float(r0["fx"]) - float(g["it0_ref_evaluate_fx"])) < 1e-10 * float(g["it0_ref_evaluate_fx"])
    assert rel_err(r0["g"], g["it0_ref_evaluate_g"])[0] < 1e-10
    assert np.array_equal(r0["g"], r0["g_again"]) and r0["fx"] == r0["fx_again"]
    for k in ("theta1", "theta1_p2p"):
        assert np.abs(r0[k] - g["ref_theta_iter1"]).max() < 1e-9


def test_one_process_driving_two_devices():
    """One process, two contexts on two GPUs (the thread-per-GPU deployment of INTEGRATION.md section 4): kernel attributes
    such as the > 48 KB dynamic shared memory opt-in are per device and must be set on each."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    from conftest import load_synth, rel_err
    pkg = load_package()
    for name in ("mlp64", "arm_sigma", "acts5"):
        s = load_synth(name)
        outs = []
        for dev in (0, 1):
            for path in (pkg.api.PATH_AUTO, pkg.api.PATH_GEMM_CHAIN):
                with pkg.Context(s["layers"], s["acfunc"], device=dev) as ctx:
                    ctx.set_path(path)
                    ctx.set_model(s["theta"])
                    ctx.set_batch(s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
                    z = ctx.fvp(s["v"], 0.1)
                    u, _ = ctx.update(0.1)
                assert rel_err(z, s["ref_fvpfast"])[0] < 1e-10, (name, dev, path)
                assert rel_err(u, s["ref_update"])[0] < 1e-8, (name, dev, path)
                outs.append((path, z))
        assert np.array_equal(outs[0][1], outs[2][1]) and np.array_equal(outs[1][1], outs[3][1])   # same bits on both GPUs
