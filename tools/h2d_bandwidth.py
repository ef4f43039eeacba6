"""Pinned host -> device copy rate of this box for a buffer of the Humanoid-size batch (1 M x 376 doubles = 3.0 GB): the floor
under the first FVP of an end-to-end solve, whose rows cannot be used before they arrive.

    python tools/h2d_bandwidth.py [bytes]"""
import json
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000 * 376 * 8
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {"bytes": n}
for pieces in (1, 12):
    step = (n + pieces - 1) // pieces
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for o in range(0, n, step):
            dst[o:o + step].copy_(src[o:o + step], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[f"ms_{pieces}_pieces"] = best
    out[f"GBps_{pieces}_pieces"] = n / best / 1e6
print(json.dumps(out))
