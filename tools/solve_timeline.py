"""Where a CG iteration of the persistent solve kernel spends its time (CTA 0's %globaltimer stamps, trpo_ctx_solve_timeline).

    python tools/solve_timeline.py [workload] [n_states]                       # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/solve_timeline.py [workload] [n_states]   # sharded over N

Prints per rank the mean microseconds of: pass (FVP over the shard), wait for the slowest CTA, slice column sums (+ push),
wait for the peers' slices, p.z round, r.r round, publication of the new direction; and the whole iteration."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import make_workload, shard_bounds  # noqa: E402

pkg = load_package()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
name = sys.argv[1] if len(sys.argv) > 1 else "mlp64"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
layers, ac, n_total, theta, batch, vec = make_workload(pkg, name, n)
lo, hi = shard_bounds(n_total, world, rank)
ctx = pkg.Context(layers, ac, device=local)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pkg.api.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.init_comm(bytes(uid.cpu().numpy().tobytes()), rank, world)
    mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).to(dev)
    allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine)
    ctx.p2p_attach(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    dist.barrier()
ctx.set_model(theta)
ctx.set_batch(np.ascontiguousarray(batch["Observ"][lo:hi]), batch["Std"])
d_b = torch.from_numpy(vec["b"]).to(dev)
d_x = torch.zeros_like(d_b)
ctx.solve_timeline(10)
acc = []
for rep in range(8):
    if world > 1:
        dist.barrier()
    ctx.cg_device(d_b.data_ptr(), d_x.data_ptr(), 10, 0.0, 0.1)
    ctx.sync()
    t = ctx.solve_timeline(10, read=True).astype(np.float64)
    if rep >= 3:
        acc.append(t)
assert ctx.solve_kernel_used()
t = np.mean(acc, axis=0) / 1e3                                    # microseconds
names = ["pass", "wait_ctas", "slice_sums", "wait_peers", "pz_round", "rr_round", "publish_p"]
d = {nm: float(np.mean(t[1:, i + 1] - t[1:, i])) for i, nm in enumerate(names)}
d["iteration"] = float(np.mean(t[1:, 7] - t[1:, 0]))
d["gap_to_next_pass"] = float(np.mean(t[2:, 0] - t[1:-1, 7]))
d["solve_us"] = float(t[-1, 7] - t[0, 0])
print(json.dumps({"rank": rank, "world": world, "workload": name, "states_per_gpu": hi - lo, **{k: round(v, 2) for k, v in d.items()}}), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
