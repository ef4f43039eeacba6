"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libtrpo_ref.so, compiled from
/root/reference by oracle/Makefile) on the ArmTest vectors and on small synthetic cases.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
The .npz files hold inputs (float64, exactly what the text files parse to) and the reference's outputs.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from oracle_lib import Oracle, Reference, build_oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
REF_BUILD = "/root/reference/build"

# small synthetic cases: (name, layers, acfunc, N, seed offset) -- exercise sigma != 1, B != 0, 'o'/'s', NumLayers != 4
SYNTH_CASES = [
    ("arm_sigma", [15, 16, 16, 3], "lttl", 500, 11),
    ("acts5", [6, 8, 7, 5, 2], "lstso", 300, 12),
    ("net3", [4, 5, 2], "ltl", 200, 13),
    ("mlp64", [17, 64, 64, 6], "lttl", 256, 14),
    ("odd_tanh_out", [11, 13, 9, 4], "ltst", 333, 15),
    ("pendulum64", [4, 64, 64, 1], "lttl", 130, 16),
]


def main():
    build_oracle()
    ref, orc = Reference(), Oracle()
    layers, ac, N = [15, 16, 16, 3], "lttl", 3150
    M, D = f"{REF_BUILD}/ArmTestModel.txt", f"{REF_BUILD}/ArmTestData.txt"
    theta = orc.load_model(M, layers, ac)
    d = orc.load_data(D, layers, ac, N)
    fvp_in, fvp_stale = pkg.textio.read_vector_pairs(f"{REF_BUILD}/ArmTestFVP.txt")
    cg_b, cg_expected = pkg.textio.read_vector_pairs(f"{REF_BUILD}/ArmTestCG.txt")
    upd_expected = pkg.textio.read_model(f"{REF_BUILD}/ArmTestModelUpdated.txt", theta.size)
    out = dict(theta=theta, fvp_in=fvp_in, fvp_expected_stale=fvp_stale, cg_b=cg_b, cg_expected=cg_expected,
               updated_expected=upd_expected, **d)
    out["ref_fvpfast_3150"], _ = ref.fvp_fast(M, D, layers, ac, 3150, 0.1, fvp_in)
    out["ref_fvpfast_2400"], _ = ref.fvp_fast(M, D, layers, ac, 2400, 0.1, fvp_in)
    out["ref_fvp4_3150"], _ = ref.fvp(M, D, layers, ac, 3150, 0.1, fvp_in)
    out["ref_fvpfast_cgb_3150"], _ = ref.fvp_fast(M, D, layers, ac, 3150, 0.1, cg_b)
    out["ref_cg_3150"], _ = ref.cg(M, D, layers, ac, 3150, 0.1, cg_b)
    out["ref_cg_2400"], _ = ref.cg(M, D, layers, ac, 2400, 0.1, cg_b)
    out["ref_update_3150"], _ = ref.update(M, D, layers, ac, 3150, 0.1)
    # the reference only printf's its CG trace; the restatement (bit-exact to it above) records the same numbers
    x, nf, rd, xn = orc.cg(layers, ac, theta, d["Std"], d["Observ"], 0.1, cg_b)
    assert np.array_equal(x, out["ref_cg_3150"])
    out["cg_trace_rdotr_3150"], out["cg_trace_xnorm_3150"] = rd, xn
    np.savez_compressed(os.path.join(HERE, "armtest.npz"), **out)

    with tempfile.TemporaryDirectory() as tmp:
        for name, layers, ac, N, off in SYNTH_CASES:
            seed = pkg.synth.SEED_BASE + off
            theta = pkg.synth.make_model(layers, seed)
            b = pkg.synth.make_batch(layers, ac, theta, N, seed)
            b["Mean"] = orc.forward(layers, ac, theta, b["Observ"])   # what the reference's own check expects
            vec = pkg.synth.make_vectors(layers, seed)
            mf, df = os.path.join(tmp, name + "_model.txt"), os.path.join(tmp, name + "_data.txt")
            pkg.textio.write_model(mf, theta)
            pkg.textio.write_data(df, b["Mean"], b["Std"], b["Observ"], b["Action"], b["Advantage"])
            o = dict(layers=np.array(layers), acfunc=np.array(ac), theta=theta, v=vec["v"], b=vec["b"], **b)
            o["ref_fvpfast"], _ = ref.fvp_fast(mf, df, layers, ac, N, 0.1, vec["v"])
            o["ref_fvp4"], _ = ref.fvp(mf, df, layers, ac, N, 0.1, vec["v"])
            o["ref_cg"], _ = ref.cg(mf, df, layers, ac, N, 0.1, vec["b"])
            o["ref_update"], _ = ref.update(mf, df, layers, ac, N, 0.1)
            np.savez_compressed(os.path.join(HERE, f"synth_{name}.npz"), **o)
            print("wrote", name)


if __name__ == "__main__":
    main()
