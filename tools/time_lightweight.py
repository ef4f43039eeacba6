"""TRPO_Lightweight (unmodified reference, oracle/_ref, one host core) against TRPO_Lightweight_GPU on the same model /
baseline files: seconds per iteration at the reference's own batch (20 episodes x 150 steps) and, GPU only, at 1 000
episodes x 1 000 steps through TRPO_Lightweight_GPU_ex. Prints one JSON line.
    python tools/time_lightweight.py [iters]"""
import contextlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lightweight_loop as lw  # noqa: E402
import oracle_lib  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "lightweight.npz")))
os.environ["TRPO_LBFGS_LIB"] = os.path.join(oracle_lib.ORACLE_DIR, "_ref", "libtrpo_ref.so")


@contextlib.contextmanager
def quiet():
    devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    pkg.textio.write_model("m.txt", gold["theta0"])
    pkg.textio.write_model("b.txt", gold["x_base0"])
    with quiet():
        t_ref = oracle_lib.Reference(fast=True).lightweight("m.txt", "b.txt", "ref", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, iters)
        pkg.api.TRPO_Lightweight_GPU("m.txt", "b.txt", "w", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, 1)          # warm-up
        t_gpu = pkg.api.TRPO_Lightweight_GPU("m.txt", "b.txt", "gpu", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, iters)
        t_big = pkg.api.TRPO_Lightweight_GPU("m.txt", "b.txt", "big", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, iters,
                                             num_ep=1000, ep_len=1000)
        # the loop is chaotic (sampling, line-search and L-BFGS branches): two CPU builds of the SAME reference source --
        # strict IEEE and -O3 with FMA contraction -- are the yardstick for what "equal" can mean after `iters` iterations
        oracle_lib.Reference(fast=False).lightweight("m.txt", "b.txt", "str", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, iters)
    ref = np.loadtxt("ref%03d.txt" % (iters - 1))
    strict = np.loadtxt("str%03d.txt" % (iters - 1))
    got = np.loadtxt("gpu%03d.txt" % (iters - 1))
print(json.dumps({"iters": iters, "reference_s_per_iter_3000_steps": t_ref / iters, "gpu_s_per_iter_3000_steps": t_gpu / iters,
                  "gpu_s_per_iter_1M_steps": t_big / iters,
                  "result_files_max_abs_diff": {"gpu_vs_reference_strict_build": float(np.abs(strict - got).max()),
                                                "gpu_vs_reference_O3_fma_build": float(np.abs(ref - got).max()),
                                                "reference_strict_vs_O3_fma_build": float(np.abs(ref - strict).max())},
                  "cores": 1}))
