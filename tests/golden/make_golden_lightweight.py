"""Generate tests/golden/lightweight.npz: rows f-3 / f-4 (rollout staging, GAE, baseline objective, whole iteration).

Runs the UNMODIFIED reference's TRPO_Lightweight (oracle/_ref, compiled from /root/reference) for 1..3 iterations and
stores the parameters it writes to its result files (%.14f), next to the first iteration's rollout batch, the
reference's own ``evaluate`` output on it and the oracle's full-precision intermediate values (the oracle loop is
checked here to land on the reference's result files).

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden_lightweight.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lightweight_loop as lw  # noqa: E402
from oracle_lib import Oracle, Reference, build_oracle  # noqa: E402

REF_BUILD = "/root/reference/build"


def main():
    build_oracle()
    ref, orc = Reference(), Oracle()
    M, B = f"{REF_BUILD}/ArmTestModel.txt", f"{REF_BUILD}/ArmTestBaseline.txt"
    theta0 = orc.load_model(M, lw.ARM_LAYERS, lw.ARM_ACFUNC)
    x_base0 = np.loadtxt(B)
    out = dict(theta0=theta0, x_base0=x_base0)
    with tempfile.TemporaryDirectory(dir="/tmp") as tmp:
        pre = os.path.join(tmp, "r")                     # result prefix must stay short (TRPO_Lightweight.c:1469-1473)
        assert len(pre) <= 22, pre
        for iters in (1, 2, 3):
            ref.lightweight(M, B, pre, lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, iters)
            out[f"ref_theta_iter{iters}"] = np.loadtxt(pre + "%03d.txt" % (iters - 1))
    trace = []
    got = lw.run(lw.OracleBackend(orc), orc, ref, theta0, x_base0, 3, trace=trace)
    assert np.abs(got - out["ref_theta_iter3"]).max() < 1e-13
    t0 = trace[0]
    for k, v in t0["batch"].items():
        out["it0_" + k] = v
    out["it0_Return"], out["it0_Advantage"] = t0["ret"], t0["adv"]
    out["it0_x_base_fitted"], out["it0_theta"] = t0["x_base"], t0["theta"]
    x = np.zeros(lw.PADDED)
    x[:x_base0.size] = x_base0
    fx, g, pred = ref.vf_evaluate(lw.ARM_VF_LAYERS, lw.ARM_ACFUNC, x, t0["batch"]["Observ"], t0["ret"], lw.NUM_EP, lw.EP_LEN)
    out["it0_ref_evaluate_fx"], out["it0_ref_evaluate_g"], out["it0_ref_evaluate_predict"] = np.float64(fx), g, pred
    for i in (1, 2):
        out[f"it{i}_theta"] = trace[i]["theta"]
    np.savez_compressed(os.path.join(HERE, "lightweight.npz"), **out)
    print("wrote lightweight.npz")


if __name__ == "__main__":
    main()
