"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the C-ABI against
 (1) the committed outputs of the compiled reference (tests/golden, FP64 tolerance 1e-10 norm-relative),
 (2) the oracle on fresh seeded inputs,
 (3) size-independent properties at BASELINE.json's full sizes (linearity, symmetry, CG residual, determinism)."""
import numpy as np
import pytest

from conftest import SYNTH_NAMES, load_synth, rel_err

pytestmark = pytest.mark.gpu

FVP_TOL = 1e-10       # north_star: FP64 build within 1e-10 relative of the reference's double-precision CPU FVP/CG
CG_TOL = 1e-8         # CG amplifies rounding by the conditioning of F + damping*I; measured values are ~1e-12
ARM_LAYERS, ARM_AC = [15, 16, 16, 3], "lttl"


def paths_for(pkg, layers, ac):
    ps = [pkg.api.PATH_GEMM_CHAIN]
    with pkg.Context(layers, ac) as c:
        try:
            c.set_path(pkg.api.PATH_FUSED)
            ps.append(pkg.api.PATH_FUSED)
        except RuntimeError:
            pass
    return ps


@pytest.mark.parametrize("n", [3150, 2400])
def test_fvp_armtest(pkg, armtest, n):
    a = armtest
    for path in paths_for(pkg, ARM_LAYERS, ARM_AC):
        with pkg.Context(ARM_LAYERS, ARM_AC) as ctx:
            ctx.set_path(path)
            ctx.set_model(a["theta"])
            ctx.set_batch(a["Observ"][:n], a["Std"])
            z = ctx.fvp(a["fvp_in"], 0.1)
            assert ctx.path_used() == path
        e_max, e_l2 = rel_err(z, a[f"ref_fvpfast_{n}"])
        assert e_max < FVP_TOL and e_l2 < FVP_TOL, (path, e_max, e_l2)
        # the 4-pass FVP() that Test_FVP_FPGA compares against (TRPOCpuCode.c:184)
        if n == 3150:
            assert rel_err(z, a["ref_fvp4_3150"])[0] < FVP_TOL


def test_cg_armtest(pkg, armtest):
    a = armtest
    for path in paths_for(pkg, ARM_LAYERS, ARM_AC):
        with pkg.Context(ARM_LAYERS, ARM_AC) as ctx:
            ctx.set_path(path)
            ctx.set_model(a["theta"])
            ctx.set_batch(a["Observ"], a["Std"])
            x, info = ctx.cg(a["cg_b"], 10, 1e-10, 0.1)
        assert info.cg_iters == 8                      # same early exit as the reference (rdotr < 1e-10 at iteration 8)
        e_max, e_l2 = rel_err(x, a["ref_cg_3150"])
        assert e_max < CG_TOL and e_l2 < CG_TOL, (path, e_max, e_l2)
        assert np.allclose(np.array(info.cg_rdotr[:8]), a["cg_trace_rdotr_3150"][:8], rtol=1e-7)
        assert np.allclose(np.array(info.cg_xnorm[:9]), a["cg_trace_xnorm_3150"][:9], rtol=1e-8)
        # the reference's own golden (ArmTestCG.txt) at its 1e-5 level
        assert rel_err(x, a["cg_expected"])[1] < 1e-5


def test_update_armtest(pkg, armtest):
    a = armtest
    with pkg.Context(ARM_LAYERS, ARM_AC) as ctx:
        ctx.set_model(a["theta"])
        ctx.set_batch(a["Observ"], a["Std"], a["Mean"], a["Action"], a["Advantage"])
        b = ctx.policy_gradient()
        u, info = ctx.update(0.1)
    assert info.ls_accepted == 1 and info.ls_steps == 1
    assert abs(info.shs - 0.00294934722) < 1e-9 and abs(info.ls_ratio[0] - 0.912970423) < 1e-7
    e_max, e_l2 = rel_err(u, a["ref_update_3150"])
    assert e_max < CG_TOL and e_l2 < CG_TOL, (e_max, e_l2)


@pytest.mark.parametrize("name", SYNTH_NAMES)
def test_synthetic_cases(pkg, oracle, name):
    s = load_synth(name)
    L, ac = s["layers"], s["acfunc"]
    for path in paths_for(pkg, L, ac):
        with pkg.Context(L, ac) as ctx:
            ctx.set_path(path)
            ctx.set_model(s["theta"])
            ctx.set_batch(s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
            z = ctx.fvp(s["v"], 0.1)
            x, info = ctx.cg(s["b"], 10, 1e-10, 0.1)
            pg = ctx.policy_gradient()
            u, uinfo = ctx.update(0.1)
        assert rel_err(z, s["ref_fvpfast"])[0] < FVP_TOL, (name, path, rel_err(z, s["ref_fvpfast"]))
        assert rel_err(z, s["ref_fvp4"])[0] < FVP_TOL
        assert rel_err(x, s["ref_cg"])[0] < CG_TOL, (name, path, rel_err(x, s["ref_cg"]))
        pg_ref = oracle.policy_gradient(L, ac, s["theta"], s["Observ"], s["Mean"], s["Action"], s["Advantage"])
        assert rel_err(pg, pg_ref)[0] < FVP_TOL
        assert rel_err(u, s["ref_update"])[0] < CG_TOL, (name, path, rel_err(u, s["ref_update"]))


def test_file_based_dropins(pkg, armtest, tmp_path, capfd):
    """FVP_GPU / CG_GPU / TRPO_Update_GPU with the reference's signature and text files (the Test_*_FPGA shape)."""
    a = armtest
    mf, df = str(tmp_path / "ArmTestModel.txt"), str(tmp_path / "ArmTestData.txt")
    pkg.textio.write_model(mf, a["theta"])
    pkg.textio.write_data(df, a["Mean"], a["Std"], a["Observ"], a["Action"], a["Advantage"])
    z, t = pkg.FVP_GPU(mf, df, ARM_LAYERS, ARM_AC, 3150, 0.1, a["fvp_in"])
    assert t >= 0 and rel_err(z, a["ref_fvpfast_3150"])[0] < FVP_TOL
    x, t = pkg.CG_GPU(mf, df, ARM_LAYERS, ARM_AC, 3150, 0.1, a["cg_b"], 10, 1e-10, 1)
    assert t >= 0 and rel_err(x, a["ref_cg_3150"])[0] < CG_TOL
    out = capfd.readouterr().out
    assert "CG Iter[0] Residual Norm=9.055952534518e-03, Soln Norm=0.000000000000e+00" in out
    assert "CG Iter[8] Residual Norm=" in out and "CG Iter[9]" not in out
    u, t = pkg.TRPO_Update_GPU(mf, df, ARM_LAYERS, ARM_AC, 3150, 0.1, 1)
    assert t >= 0 and rel_err(u, a["ref_update_3150"])[0] < CG_TOL
    out = capfd.readouterr().out
    assert "shs: 0.002949347" in out and "a/e/r 0.009916" in out


@pytest.mark.parametrize("n", [1, 7, 63, 64, 65, 1000, 4097])
def test_ragged_sample_counts(pkg, oracle, n):
    layers, ac = [17, 64, 64, 6], "lttl"
    theta = pkg.synth.make_model(layers, 99)
    batch = pkg.synth.make_batch(layers, ac, theta, n, 99)
    vec = pkg.synth.make_vectors(layers, 99)
    ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
    for path in paths_for(pkg, layers, ac):
        with pkg.Context(layers, ac) as ctx:
            ctx.set_path(path)
            ctx.set_model(theta)
            ctx.set_batch(batch["Observ"], batch["Std"])
            z = ctx.fvp(vec["v"], 0.1)
        assert rel_err(z, ref)[0] < FVP_TOL, (n, path, rel_err(z, ref))


def test_cg_termination_semantics(pkg, oracle):
    """rdotr < ResidualTh before the first FVP -> zero iterations, x = 0; MaxIter = 0 -> x = 0 (TRPO_CG.c:59-62)."""
    s = load_synth("net3")
    with pkg.Context(s["layers"], s["acfunc"]) as ctx:
        ctx.set_model(s["theta"])
        ctx.set_batch(s["Observ"], s["Std"])
        x, info = ctx.cg(s["b"], 10, 1e3, 0.1)
        assert info.cg_iters == 0 and not x.any()
        x, info = ctx.cg(s["b"], 0, 1e-10, 0.1)
        assert info.cg_iters == 0 and not x.any()
        x3, info = ctx.cg(s["b"], 3, 0.0, 0.1)
        assert info.cg_iters == 3
    ref3, nf, _, _ = oracle.cg(s["layers"], s["acfunc"], s["theta"], s["Std"], s["Observ"], 0.1, s["b"], 3, 0.0)
    assert nf == 3 and rel_err(x3, ref3)[0] < CG_TOL


def test_humanoid_width_against_oracle(pkg, oracle):
    layers, ac = [376, 256, 256, 17], "lttl"
    theta = pkg.synth.make_model(layers, 4)
    batch = pkg.synth.make_batch(layers, ac, theta, 300, 4)
    vec = pkg.synth.make_vectors(layers, 4)
    ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
    with pkg.Context(layers, ac) as ctx:
        ctx.set_model(theta)
        ctx.set_batch(batch["Observ"], batch["Std"])
        z = ctx.fvp(vec["v"], 0.1)
    assert rel_err(z, ref)[0] < FVP_TOL, rel_err(z, ref)


def test_full_size_properties_mlp64_1m(pkg, oracle):
    """BASELINE config 3 at full size (1M states): determinism, linearity, symmetry, CG residual, and a
    50k-sample prefix against the oracle."""
    layers, ac = [17, 64, 64, 6], "lttl"
    seed = pkg.synth.SEED_BASE + 2
    theta = pkg.synth.make_model(layers, seed)
    N = 1_000_000
    batch = pkg.synth.make_batch(layers, ac, theta, N, seed)
    vec = pkg.synth.make_vectors(layers, seed)
    rng = np.random.default_rng(5)
    u, w = rng.standard_normal(theta.size), rng.standard_normal(theta.size)
    with pkg.Context(layers, ac) as ctx:
        ctx.set_model(theta)
        ctx.set_batch(batch["Observ"], batch["Std"])
        Fu, Fw = ctx.fvp(u, 0.0), ctx.fvp(w, 0.0)
        assert np.array_equal(Fu, ctx.fvp(u, 0.0))                      # bitwise deterministic
        Fc = ctx.fvp(2 * u - 3 * w, 0.0)
        assert rel_err(Fc, 2 * Fu - 3 * Fw)[0] < 1e-11                   # linear
        assert abs(u @ Fw - w @ Fu) < 1e-10 * abs(u @ Fw)                # symmetric
        x, info = ctx.cg(vec["b"], 10, 0.0, 0.1)
        assert info.cg_iters == 10
        resid = ctx.fvp(x, 0.1) - vec["b"]
        assert np.linalg.norm(resid) ** 2 < 1.01 * info.cg_rdotr[10] + 1e-20   # r tracked by CG == b - A x
        # streamed staging: pinned source, the first FVP overlaps the chunked H2D copy -- bitwise the same result
        import torch
        pinned = torch.from_numpy(batch["Observ"]).pin_memory()
        for _ in range(2):
            ctx.set_batch(pinned.numpy(), batch["Std"])
            assert np.array_equal(Fu, ctx.fvp(u, 0.0))
            assert np.array_equal(Fw, ctx.fvp(w, 0.0))
        x2, _ = ctx.cg(vec["b"], 10, 0.0, 0.1)
        ctx.set_batch(pinned.numpy(), batch["Std"])
        x3, _ = ctx.cg(vec["b"], 10, 0.0, 0.1)
        assert np.array_equal(x, x2) and np.array_equal(x, x3)
        assert pkg.api.lib().trpo_ctx_comm_error(ctx.h) == 0
        # prefix vs oracle
        n = 50_000
        ctx.set_batch(batch["Observ"][:n], batch["Std"])
        z = ctx.fvp(vec["v"], 0.1)
    ref = oracle.fvp(layers, ac, theta, batch["Std"], np.ascontiguousarray(batch["Observ"][:n]), 0.1, vec["v"])
    assert rel_err(z, ref)[0] < FVP_TOL, rel_err(z, ref)


def test_c_harness_binary(pkg, armtest, tmp_path):
    """The plain-C driver (host/trpo_test_main.c, the shape of the reference's Test_CG / Test_FVP_FPGA harness,
    TRPOCpuCode.c:76-223) linked against libtrpo_b200.so: no Python between the caller and the C-ABI."""
    import os
    import re
    import subprocess
    a = armtest
    exe = os.path.join(os.path.dirname(pkg.api.library_path()), "trpo_test_gpu")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    mf, df = str(tmp_path / "ArmTestModel.txt"), str(tmp_path / "ArmTestData.txt")
    pkg.textio.write_model(mf, a["theta"])
    pkg.textio.write_data(df, a["Mean"], a["Std"], a["Observ"], a["Action"], a["Advantage"])
    # vectors file in the ArmTestCG.txt format: "input expected" with the compiled reference's result as expectation
    cgf = str(tmp_path / "cg.txt")
    with open(cgf, "w") as f:
        for b, e in zip(a["cg_b"], a["ref_cg_3150"]):
            f.write(f"{float(b)!r} {float(e)!r}\n")
    out = subprocess.run([exe, "cg", mf, df, "3150", cgf], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "CG Iter[8] Residual Norm=" in out.stdout
    m = re.search(r"max\|d\|/max\|ref\| = ([0-9.e+-]+), rel-L2 = ([0-9.e+-]+)", out.stdout)
    assert m and float(m.group(1)) < CG_TOL and float(m.group(2)) < CG_TOL, out.stdout[-400:]
    fvf = str(tmp_path / "fvp.txt")
    with open(fvf, "w") as f:
        for b, e in zip(a["fvp_in"], a["ref_fvpfast_3150"]):
            f.write(f"{float(b)!r} {float(e)!r}\n")
    out = subprocess.run([exe, "fvp", mf, df, "3150", fvf], capture_output=True, text=True, timeout=300)
    m = re.search(r"max\|d\|/max\|ref\| = ([0-9.e+-]+), rel-L2 = ([0-9.e+-]+)", out.stdout)
    assert out.returncode == 0 and m and float(m.group(1)) < FVP_TOL, out.stdout[-400:]
    # a missing model file: the reference's error text and a non-zero exit
    out = subprocess.run([exe, "fvp", str(tmp_path / "none.txt"), df, "3150", fvf], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "[ERROR] Cannot open Model File" in out.stderr


def test_reference_symbol_names_in_dropin_library(pkg, armtest, tmp_path):
    """FVP_FPGA / CG_FPGA (TRPO.h:98,101) resolved from libtrpo_b200_dropin.so and called with TRPOparam by value."""
    import ctypes as C
    a = armtest
    mf, df = str(tmp_path / "m.txt"), str(tmp_path / "d.txt")
    pkg.textio.write_model(mf, a["theta"])
    pkg.textio.write_data(df, a["Mean"], a["Std"], a["Observ"], a["Action"], a["Advantage"])
    lib = C.CDLL(pkg.api.library_path(dropin=True))
    lib.FVP_FPGA.restype = C.c_double
    lib.FVP_FPGA.argtypes = [pkg.TRPOparam, pkg.api.c_double_p, pkg.api.c_double_p]
    lib.CG_FPGA.restype = C.c_double
    lib.CG_FPGA.argtypes = [pkg.TRPOparam, pkg.api.c_double_p, pkg.api.c_double_p, C.c_size_t, C.c_double, C.c_size_t]
    keep = []
    p = pkg.api.make_param(mf, df, ARM_LAYERS, ARM_AC, 3150, 0.1, keep)
    out = np.zeros(582)
    v = np.ascontiguousarray(a["fvp_in"])
    t = lib.FVP_FPGA(p, out.ctypes.data_as(pkg.api.c_double_p), v.ctypes.data_as(pkg.api.c_double_p))
    assert t >= 0 and rel_err(out, a["ref_fvpfast_3150"])[0] < FVP_TOL
    b = np.ascontiguousarray(a["cg_b"])
    t = lib.CG_FPGA(p, out.ctypes.data_as(pkg.api.c_double_p), b.ctypes.data_as(pkg.api.c_double_p), 10, 1e-10, 4)
    assert t >= 0 and rel_err(out, a["ref_cg_3150"])[0] < CG_TOL


FP32_TOL = 1e-4       # stated tolerance of the optional FP32 mode (north_star: "~1e-4")


@pytest.mark.parametrize("name", ["mlp64", "arm_sigma", "acts5", "odd_tanh_out"])
def test_fp32_mode_within_stated_tolerance(pkg, name):
    """Optional FP32 mode (3xTF32 tensor-core products, FP32 slice sums, FP64 reduction + CG) against the compiled
    reference's FP64 results: norm-relative error below the stated 1e-4."""
    s = load_synth(name)
    L, ac = s["layers"], s["acfunc"]
    with pkg.Context(L, ac, precision=pkg.api.PRECISION_FP32) as ctx:
        ctx.set_model(s["theta"])
        ctx.set_batch(s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
        z = ctx.fvp(s["v"], 0.1)
        x, info = ctx.cg(s["b"], 10, 1e-10, 0.1)
    e = rel_err(z, s["ref_fvpfast"])
    assert e[0] < FP32_TOL and e[1] < FP32_TOL, (name, e)
    assert e[1] > 1e-12                       # it really is the FP32 path
    e = rel_err(x, s["ref_cg"])
    assert e[1] < 5e-3, (name, e)             # CG amplifies the FVP error by the conditioning of F + damping*I


def test_fp32_mode_humanoid_width(pkg, oracle):
    layers, ac = [376, 256, 256, 17], "lttl"
    theta = pkg.synth.make_model(layers, 4)
    batch = pkg.synth.make_batch(layers, ac, theta, 300, 4)
    vec = pkg.synth.make_vectors(layers, 4)
    ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
    with pkg.Context(layers, ac, precision=pkg.api.PRECISION_FP32) as ctx:
        ctx.set_model(theta)
        ctx.set_batch(batch["Observ"], batch["Std"])
        z = ctx.fvp(vec["v"], 0.1)
    e = rel_err(z, ref)
    assert e[0] < FP32_TOL and e[1] < FP32_TOL, e


@pytest.mark.parametrize("layers,ac", [([5, 3], "ll"), ([9, 4], "lt"), ([6, 7, 8, 9, 10, 3], "ltstol"),
                                       ([20, 64, 64, 8], "lttl"), ([3, 130, 70, 2], "ltsl"),
                                       ([11, 32, 32, 3], "lttl"), ([32, 24, 32, 8], "lsto"), ([16, 16, 16, 8], "ltto"),
                                       # widths that are whole tiles / whole k-steps: bias as accumulator start value
                                       # (forward) and as column sums (outer product) instead of an augmented row
                                       ([128, 256, 128, 5], "ltsl"), ([32, 128, 3], "ltl")])
def test_unusual_depths_and_widths(pkg, oracle, layers, ac):
    """NumLayers 2 and 6, widths that straddle the tile sizes, shapes at the fused kernel's eligibility limits."""
    seed = 1000 + sum(layers)
    theta = pkg.synth.make_model(layers, seed)
    batch = pkg.synth.make_batch(layers, ac, theta, 777, seed)
    batch["Mean"] = oracle.forward(layers, ac, theta, batch["Observ"])
    vec = pkg.synth.make_vectors(layers, seed)
    z_ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
    x_ref, nf, _, _ = oracle.cg(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["b"])
    u_ref, uinfo = oracle.update(layers, ac, theta, batch["Std"], batch["Observ"], batch["Mean"], batch["Action"],
                                 batch["Advantage"], 0.1)
    for path in paths_for(pkg, layers, ac):
        with pkg.Context(layers, ac) as ctx:
            ctx.set_path(path)
            ctx.set_model(theta)
            ctx.set_batch(batch["Observ"], batch["Std"], batch["Mean"], batch["Action"], batch["Advantage"])
            z = ctx.fvp(vec["v"], 0.1)
            x, info = ctx.cg(vec["b"], 10, 1e-10, 0.1)
            u, ginfo = ctx.update(0.1)
        assert rel_err(z, z_ref)[0] < FVP_TOL, (layers, path, rel_err(z, z_ref))
        # 10 CG iterations amplify the 1e-14 summation-order differences of the FVP by the conditioning of F + 0.1 I:
        # with far more parameters than samples (3-130-70-2: P = 9 832, N = 777) the two FP64 trajectories agree to
        # ~1e-6 only, which is a property of the problem, not of the kernels (the FVP itself is at 1e-14)
        cg_tol = CG_TOL if theta.size < 777 else 1e-4
        assert info.cg_iters == nf and rel_err(x, x_ref)[0] < cg_tol, (layers, path, rel_err(x, x_ref))
        assert ginfo.ls_steps == uinfo.ls_steps and ginfo.ls_accepted == uinfo.ls_accepted
        assert rel_err(u, u_ref)[0] < cg_tol, (layers, path, rel_err(u, u_ref))


def test_repeated_solves_replay_a_cuda_graph_bitwise(pkg, armtest):
    """The 2nd identical solve is captured into a CUDA graph and the following ones replay it: every repetition must
    return bitwise the same x and the same trace as the first (directly launched) one; a different b falls back."""
    a = armtest
    with pkg.Context(ARM_LAYERS, ARM_AC) as ctx:
        ctx.set_model(a["theta"])
        ctx.set_batch(a["Observ"], a["Std"])
        xs, iters = [], []
        for _ in range(5):
            x, info = ctx.cg(a["cg_b"], 10, 1e-10, 0.1)
            xs.append(x); iters.append(info.cg_iters)
        x_other, _ = ctx.cg(2.0 * a["cg_b"], 10, 1e-10, 0.1)
        x_back, _ = ctx.cg(a["cg_b"], 10, 1e-10, 0.1)
        # a new model must be picked up by the replayed graph (same device pointers, new contents)
        ctx.set_model(1.01 * a["theta"])
        x_new_model, _ = ctx.cg(a["cg_b"], 10, 1e-10, 0.1)
    assert all(np.array_equal(xs[0], x) for x in xs[1:]) and iters == [8] * 5
    assert np.array_equal(xs[0], x_back)
    assert rel_err(xs[0], a["ref_cg_3150"])[0] < CG_TOL
    assert rel_err(x_other, 2.0 * a["ref_cg_3150"])[0] < 1e-6       # linear system: x scales with b (different early exit point)
    assert not np.array_equal(x_new_model, xs[0])


@pytest.mark.parametrize("variant", ["8", "4"])
def test_block_wide_variants_of_the_arm_kernel(variant, tmp_path):
    """TRPO_FUSED_ARM_VARIANT selects the block-wide fused kernel (one 8-warp CTA, or four 4-warp CTAs per SM) instead of
    the warp-private default for the 15-16-16-3 policy; the switch is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from __graft_entry__ import load_package\n"
        "pkg = load_package()\n"
        f"a = dict(np.load({os.path.join(root, 'tests', 'golden', 'armtest.npz')!r}))\n"
        "with pkg.Context([15, 16, 16, 3], 'lttl') as ctx:\n"
        "    ctx.set_model(a['theta']); ctx.set_batch(a['Observ'], a['Std'])\n"
        "    z = ctx.fvp(a['fvp_in'], 0.1); x, info = ctx.cg(a['cg_b'], 10, 1e-10, 0.1)\n"
        "    assert ctx.path_used() == 2\n"
        "e1 = np.abs(z - a['ref_fvpfast_3150']).max() / np.abs(a['ref_fvpfast_3150']).max()\n"
        "e2 = np.abs(x - a['ref_cg_3150']).max() / np.abs(a['ref_cg_3150']).max()\n"
        "print('ERR', e1, e2, info.cg_iters)\n"
        "assert e1 < 1e-10 and e2 < 1e-8 and info.cg_iters == 8\n"
    )
    env = dict(os.environ, TRPO_FUSED_ARM_VARIANT=variant)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-1500:]


def test_update_with_failed_line_search_returns_the_step_direction(pkg, oracle):
    """TRPO_Update.c:852 quirk: when no line-search step is accepted the step direction is returned, not the parameters.
    Zero advantages give b = 0: CG stops before its first FVP and the direction is the zero vector on both sides."""
    s = load_synth("net3")
    L, ac = s["layers"], s["acfunc"]
    zero_adv = np.zeros_like(s["Advantage"])
    u_ref, info_ref = oracle.update(L, ac, s["theta"], s["Std"], s["Observ"], s["Mean"], s["Action"], zero_adv, 0.1)
    with pkg.Context(L, ac) as ctx:
        ctx.set_model(s["theta"])
        ctx.set_batch(s["Observ"], s["Std"], s["Mean"], s["Action"], zero_adv)
        u, info = ctx.update(0.1)
    assert info.cg_iters == info_ref.cg_iters == 0
    assert info.ls_accepted == info_ref.ls_accepted == 0 and info.ls_steps == info_ref.ls_steps == 10
    assert np.array_equal(np.nan_to_num(u), np.nan_to_num(u_ref)) and not np.nan_to_num(u).any()


@pytest.mark.parametrize("name", ["acts5", "mlp64", "odd_tanh_out"])
def test_forward_pass_matches_oracle(pkg, oracle, name):
    """trpo_ctx_forward: the policy mean of every staged sample (TRPO_Update.c:259-291)."""
    s = load_synth(name)
    with pkg.Context(s["layers"], s["acfunc"]) as ctx:
        ctx.set_model(s["theta"])
        ctx.set_batch(s["Observ"], s["Std"])
        mean = ctx.forward(s["Observ"].shape[0])
    ref = oracle.forward(s["layers"], s["acfunc"], s["theta"], s["Observ"])
    assert np.abs(mean - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------------------------------------------------------------
# round 2: the parity holes the round-1 review named

@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("layers,ac,n,chunk", [([17, 64, 64, 6], "lttl", 2500, 1024),        # 3 passes, ragged last (452)
                                               ([376, 256, 256, 17], "lttl", 1100, 384),     # Humanoid width: 384+384+332
                                               ([6, 7, 8, 9, 10, 3], "ltstol", 1000, 256),   # 4 passes, 6 layers, no tail kernel
                                               ([9, 40, 30], "ltt", 700, 128)])              # tanh output layer: Y_K is kept
def test_multi_chunk_gemm_chain(pkg, oracle, layers, ac, n, chunk, precision):
    """A batch larger than the GEMM-chain chunk runs as several passes whose per-slice partial sums accumulate in place
    (chain_accumulate's `accumulate = chunk_idx > 0`; the sum the reference forms sample by sample, TRPO_FVP.c:903-921).
    trpo_ctx_set_chunk forces >= 3 passes with a ragged last one: FVP, policy gradient and the whole update against the
    oracle, and against the same context's single-pass result."""
    seed = 500 + sum(layers)
    theta = pkg.synth.make_model(layers, seed)
    batch = pkg.synth.make_batch(layers, ac, theta, n, seed)
    batch["Mean"] = oracle.forward(layers, ac, theta, batch["Observ"])
    vec = pkg.synth.make_vectors(layers, seed)
    z_ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
    pg_ref = oracle.policy_gradient(layers, ac, theta, batch["Observ"], batch["Mean"], batch["Action"], batch["Advantage"])
    u_ref, uinfo = oracle.update(layers, ac, theta, batch["Std"], batch["Observ"], batch["Mean"], batch["Action"],
                                 batch["Advantage"], 0.1)
    fp32 = precision == "fp32"
    out = {}
    for tag, ch in (("multi", chunk), ("single", 0)):
        with pkg.Context(layers, ac, precision=pkg.api.PRECISION_FP32 if fp32 else pkg.api.PRECISION_FP64) as ctx:
            ctx.set_path(pkg.api.PATH_GEMM_CHAIN)
            ctx.set_chunk(ch)
            ctx.set_model(theta)
            ctx.set_batch(batch["Observ"], batch["Std"], batch["Mean"], batch["Action"], batch["Advantage"])
            z = ctx.fvp(vec["v"], 0.1)
            used = ctx.chunk_used()
            pg = ctx.policy_gradient()
            u, ginfo = ctx.update(0.1)
            out[tag] = (z, pg, u, ginfo.ls_steps, ginfo.ls_accepted)
        if tag == "multi":
            assert used == (chunk + 127) // 128 * 128 and (n + used - 1) // used >= 3, (used, n)
        else:
            assert used >= n
    fvp_tol = FP32_TOL if fp32 else FVP_TOL
    for tag in out:
        z, pg, u, steps, acc = out[tag]
        e = rel_err(z, z_ref)
        assert e[0] < fvp_tol and e[1] < fvp_tol, (tag, layers, precision, e)
        assert rel_err(pg, pg_ref)[0] < FVP_TOL, (tag, rel_err(pg, pg_ref))          # the policy gradient is FP64 in both modes
        if not fp32:
            cg_tol = CG_TOL if theta.size < n else 1e-4                              # see test_unusual_depths_and_widths
            assert steps == uinfo.ls_steps and acc == uinfo.ls_accepted
            assert rel_err(u, u_ref)[0] < cg_tol, (tag, layers, rel_err(u, u_ref))
        else:
            assert np.isfinite(u).all()
    # several passes only change the order in which slices receive their samples
    assert rel_err(out["multi"][0], out["single"][0])[0] < (FP32_TOL if fp32 else 1e-12)
    assert rel_err(out["multi"][1], out["single"][1])[0] < 1e-12


def test_batch_without_mean_action_advantage_invalidates_the_old_ones(pkg):
    """A second trpo_ctx_set_batch WITHOUT Mean/Action/Advantage must not leave the previous (smaller) arrays paired with
    the new rows: the policy gradient / update then fail instead of reading stale or out-of-bounds memory."""
    s = load_synth("mlp64")
    L, ac = s["layers"], s["acfunc"]
    n = s["Observ"].shape[0]
    with pkg.Context(L, ac) as ctx:
        ctx.set_model(s["theta"])
        ctx.set_batch(s["Observ"][: n // 2], s["Std"], s["Mean"][: n // 2], s["Action"][: n // 2], s["Advantage"][: n // 2])
        ctx.policy_gradient()
        ctx.set_batch(s["Observ"], s["Std"])                       # more rows, observations only
        assert rel_err(ctx.fvp(s["v"], 0.1), s["ref_fvpfast"])[0] < FVP_TOL
        with pytest.raises(RuntimeError, match="Mean/Action/Advantage"):
            ctx.policy_gradient()
        with pytest.raises(RuntimeError, match="Mean/Action/Advantage"):
            ctx.update(0.1)
        ctx.set_batch(s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
        u, _ = ctx.update(0.1)
    assert rel_err(u, s["ref_update"])[0] < CG_TOL


def test_max_iter_is_unbounded_like_the_reference(pkg, oracle, armtest, tmp_path, capfd):
    """TRPO_CG.c:45 loops `for iter = 0 .. MaxIter` with no upper bound on MaxIter; round 1 refused MaxIter > 32."""
    a = armtest
    with pkg.Context(ARM_LAYERS, ARM_AC) as ctx:
        ctx.set_model(a["theta"])
        ctx.set_batch(a["Observ"], a["Std"])
        x, info = ctx.cg(a["cg_b"], 1000, 1e-10, 0.1)               # early exit at iteration 8, as with MaxIter = 10
        assert info.cg_iters == 8 and rel_err(x, a["ref_cg_3150"])[0] < CG_TOL
        x40, info = ctx.cg(a["cg_b"], 40, 0.0, 0.1)                 # ResidualTh = 0: all 40 FVPs run
        assert info.cg_iters == 40
        rd, xn = ctx.cg_trace()
        assert len(rd) == 41 and len(xn) == 41 and np.isfinite(rd).all() and np.isfinite(xn).all()
        assert np.allclose(rd[:8], a["cg_trace_rdotr_3150"][:8], rtol=1e-7)
        assert np.array_equal(np.array(info.cg_rdotr[:34]), rd[:34])
    ref40, nf, _, _ = oracle.cg(ARM_LAYERS, ARM_AC, a["theta"], a["Std"], a["Observ"], 0.1, a["cg_b"], 40, 0.0)
    assert nf == 40 and rel_err(x40, ref40)[0] < 1e-6              # 30 further iterations on a converged system
    mf, df = str(tmp_path / "m.txt"), str(tmp_path / "d.txt")
    pkg.textio.write_model(mf, a["theta"])
    pkg.textio.write_data(df, a["Mean"], a["Std"], a["Observ"], a["Action"], a["Advantage"])
    capfd.readouterr()
    x, t = pkg.CG_GPU(mf, df, ARM_LAYERS, ARM_AC, 3150, 0.1, a["cg_b"], 40, 0.0, 1)
    out = capfd.readouterr().out
    assert t >= 0 and "CG Iter[40] Residual Norm=" in out and "CG Iter[41]" not in out


def test_reference_own_harness_against_the_dropin(pkg, armtest, tmp_path):
    """The reference's OWN Test_FVP_FPGA / Test_CG_FPGA (/root/reference/src/TRPOCpuCode.c:138-311), compiled unmodified by
    oracle/Makefile into oracle/_ref/ref_harness and linked against libtrpo_b200_dropin.so in the FPGA library's place: it
    runs the reference's CPU FVP() / CG() beside FVP_FPGA / CG_FPGA on the ArmTest files and prints its MAPE."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "ref_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_harness was not built (needs /root/reference at build time)")
    a = armtest
    pkg.textio.write_model(str(tmp_path / "ArmTestModel.txt"), a["theta"])
    pkg.textio.write_data(str(tmp_path / "ArmTestData.txt"), a["Mean"], a["Std"], a["Observ"], a["Action"], a["Advantage"])
    for fname, vin, vexp in (("ArmTestFVP.txt", a["fvp_in"], a["ref_fvpfast_3150"]), ("ArmTestCG.txt", a["cg_b"], a["cg_expected"])):
        with open(tmp_path / fname, "w") as f:
            for x, e in zip(vin, vexp):
                f.write(f"{float(x)!r} {float(e)!r}\n")
    out = subprocess.run([exe, "all"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-800:]
    assert "[ERROR]" not in out.stderr, out.stderr[-800:]
    m_fvp = re.search(r"Test FPGA -+\n\[INFO\] FPGA Computing Time = ([0-9.e+-]+) seconds\n\[INFO\] Mean Absolute Percentage Error = ([0-9.e+-]+)%", out.stdout)
    m_cg = re.search(r"Mean Absolute Percentage Error = ([0-9.e+-]+)%, Max Percentage Error = ([0-9.e+-]+)%", out.stdout)
    assert m_fvp and m_cg, out.stdout[-1500:]
    # the harness's own metric, in percent: FVP elements agree to ~1e-12 %, CG (8 iterations) to ~1e-8 %
    assert float(m_fvp.group(2)) < 1e-8, out.stdout[-1500:]
    assert float(m_cg.group(1)) < 1e-6 and float(m_cg.group(2)) < 1e-4, out.stdout[-1500:]
    assert "Difference" not in out.stdout                           # no element off by more than 1 %
    # both CG traces were printed: the GPU's (from CG_FPGA) and the reference's (from CG)
    assert out.stdout.count("CG Iter[8] Residual Norm=") == 2


def _run_with_env(code, env_extra, timeout=600):
    """Kernel-selection switches are read once per process: run a snippet in a subprocess with the given environment."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prelude = ("import sys, json, numpy as np\n"
               f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
               "from __graft_entry__ import load_package\n"
               "from conftest import load_synth\n"
               "pkg = load_package()\n")
    out = subprocess.run([sys.executable, "-c", prelude + code], env=dict(os.environ, **env_extra), capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0, out.stdout[-800:] + out.stderr[-2000:]
    return out.stdout


SOLVE_SNIPPET = (
    "res = {}\n"
    "for name in ('mlp64', 'arm_sigma', 'pendulum64'):\n"
    "    s = load_synth(name)\n"
    "    with pkg.Context(s['layers'], s['acfunc']) as ctx:\n"
    "        ctx.set_model(s['theta']); ctx.set_batch(s['Observ'], s['Std'], s['Mean'], s['Action'], s['Advantage'])\n"
    "        x, info = ctx.cg(s['b'], 10, 1e-10, 0.1)\n"
    "        x0, info0 = ctx.cg(s['b'], 10, 0.0, 0.1)\n"
    "        u, _ = ctx.update(0.1)\n"
    "        res[name] = dict(x=x.tolist(), iters=info.cg_iters, rd=list(info.cg_rdotr[:11]), x0=x0.tolist(), iters0=info0.cg_iters,\n"
    "                         u=u.tolist(), solve_kernel=ctx.solve_kernel_used(),\n"
    "                         e=float(np.abs(x - s['ref_cg']).max() / np.abs(s['ref_cg']).max()),\n"
    "                         eu=float(np.abs(u - s['ref_update']).max() / np.abs(s['ref_update']).max()))\n"
    "print('RESULT' + json.dumps(res))\n")


def test_persistent_solve_kernel_against_per_iteration_launches():
    """The whole CG as ONE cooperative kernel (in-kernel row reduction, grid barriers, CG update) and as per-iteration launches
    (FVP kernel, row reduction, single-CTA update): same iteration counts, same early exit, both within tolerance of the
    compiled reference, and within rounding of each other (the dot products are summed in different fixed orders)."""
    import json
    a = json.loads(_run_with_env(SOLVE_SNIPPET, {"TRPO_FUSED_SOLVE": "1"}).split("RESULT")[1])
    b = json.loads(_run_with_env(SOLVE_SNIPPET, {"TRPO_NO_FUSED_SOLVE": "1"}).split("RESULT")[1])
    for name in a:
        assert a[name]["solve_kernel"] is True and b[name]["solve_kernel"] is False
        assert a[name]["iters"] == b[name]["iters"] and a[name]["iters0"] == b[name]["iters0"] == 10
        assert a[name]["e"] < CG_TOL and b[name]["e"] < CG_TOL and a[name]["eu"] < CG_TOL and b[name]["eu"] < CG_TOL
        xa, xb = np.array(a[name]["x"]), np.array(b[name]["x"])
        assert rel_err(xa, xb)[0] < 1e-10
        n = a[name]["iters"] + 1
        assert np.allclose(a[name]["rd"][:n], b[name]["rd"][:n], rtol=1e-9)


def test_warp_group_variants_of_the_fused_kernel():
    """TRPO_FUSED_GROUPS = 1 | 2: one 8-warp group per CTA (block barriers) or two independent 4-warp groups with the
    accumulators parked in Tensor Memory (tcgen05.st / tcgen05.ld): same results within rounding, both within tolerance."""
    import json
    snippet = (
        "res = {}\n"
        "for name in ('mlp64', 'pendulum64'):\n"
        "    s = load_synth(name)\n"
        "    with pkg.Context(s['layers'], s['acfunc']) as ctx:\n"
        "        ctx.set_model(s['theta']); ctx.set_batch(s['Observ'], s['Std'], s['Mean'], s['Action'], s['Advantage'])\n"
        "        z = ctx.fvp(s['v'], 0.1); pg = ctx.policy_gradient()\n"
        "        res[name] = dict(z=z.tolist(), pg=pg.tolist(), e=float(np.abs(z - s['ref_fvpfast']).max() / np.abs(s['ref_fvpfast']).max()))\n"
        "layers, ac = [11, 32, 32, 3], 'lttl'\n"
        "theta = pkg.synth.make_model(layers, 3); b = pkg.synth.make_batch(layers, ac, theta, 5000, 3); v = pkg.synth.make_vectors(layers, 3)\n"
        "with pkg.Context(layers, ac) as ctx:\n"
        "    ctx.set_model(theta); ctx.set_batch(b['Observ'], b['Std'], b['Mean'], b['Action'], b['Advantage'])\n"
        "    res['h32'] = dict(z=ctx.fvp(v['v'], 0.1).tolist(), pg=ctx.policy_gradient().tolist(), e=0.0)\n"
        "print('RESULT' + json.dumps(res))\n")
    one = json.loads(_run_with_env(snippet, {"TRPO_FUSED_GROUPS": "1"}).split("RESULT")[1])
    two = json.loads(_run_with_env(snippet, {"TRPO_FUSED_GROUPS": "2"}).split("RESULT")[1])
    for name in one:
        assert one[name]["e"] < FVP_TOL and two[name]["e"] < FVP_TOL
        assert rel_err(np.array(one[name]["z"]), np.array(two[name]["z"]))[0] < 1e-12
        assert rel_err(np.array(one[name]["pg"]), np.array(two[name]["pg"]))[0] < 1e-12


def test_tail_of_the_batch_is_split_in_eight_sample_units(pkg, oracle):
    """The last, partial round of tiles is dealt out in 8-sample units (one partial tile per warp group): sizes around the
    full-round boundary of 148 CTAs x 64 samples, with tails that are not multiples of 8."""
    layers, ac = [17, 64, 64, 6], "lttl"
    theta = pkg.synth.make_model(layers, 21)
    vec = pkg.synth.make_vectors(layers, 21)
    for n in (9472 - 3, 9472, 9472 + 5, 2 * 9472 + 8 * 148 + 1, 20001):
        batch = pkg.synth.make_batch(layers, ac, theta, n, 21)
        ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
        pg_ref = oracle.policy_gradient(layers, ac, theta, batch["Observ"], batch["Mean"], batch["Action"], batch["Advantage"])
        with pkg.Context(layers, ac) as ctx:
            ctx.set_model(theta)
            ctx.set_batch(batch["Observ"], batch["Std"], batch["Mean"], batch["Action"], batch["Advantage"])
            z = ctx.fvp(vec["v"], 0.1)
            pg = ctx.policy_gradient()
        assert rel_err(z, ref)[0] < FVP_TOL, (n, rel_err(z, ref))
        assert rel_err(pg, pg_ref)[0] < FVP_TOL, (n, rel_err(pg, pg_ref))


def test_fp32_tcgen05_layers_against_the_legacy_tensor_path():
    """FP32 mode at Humanoid width: the forward layers on tcgen05 (TMA-fed, TMEM accumulators) and the fused tail kernel against the
    mma.sync kernels they replace (TRPO_NO_TCGEN05 / TRPO_NO_F32_TAIL): both within the stated 1e-4 of the FP64 oracle, and
    within 3xTF32 rounding of each other; ragged row counts and a 376-wide (not a multiple of 16) first layer included."""
    import json
    snippet = (
        "sys.path.insert(0, 'tests')\n"
        "from oracle_lib import Oracle\n"
        "res = {}\n"
        "for layers, n in (([376, 256, 256, 17], 1000), ([376, 64, 64, 17], 777), ([40, 96, 32, 5], 300)):\n"
        "    ac = 'lttl'\n"
        "    theta = pkg.synth.make_model(layers, 4); b = pkg.synth.make_batch(layers, ac, theta, n, 4); v = pkg.synth.make_vectors(layers, 4)\n"
        "    ref = Oracle(fast=True).fvp(layers, ac, theta, b['Std'], b['Observ'], 0.1, v['v'])\n"
        "    with pkg.Context(layers, ac, precision=pkg.api.PRECISION_FP32) as ctx:\n"
        "        ctx.set_model(theta); ctx.set_batch(b['Observ'], b['Std'])\n"
        "        z = ctx.fvp(v['v'], 0.1)\n"
        "        ctx.set_chunk(256)\n"
        "        z2 = ctx.fvp(v['v'], 0.1)\n"
        "    res[str(layers)] = dict(z=z.tolist(), e=float(np.linalg.norm(z - ref) / np.linalg.norm(ref)),\n"
        "                            e2=float(np.linalg.norm(z2 - ref) / np.linalg.norm(ref)))\n"
        "print('RESULT' + json.dumps(res))\n")
    new = json.loads(_run_with_env(snippet, {}).split("RESULT")[1])
    old = json.loads(_run_with_env(snippet, {"TRPO_NO_TCGEN05": "1", "TRPO_NO_F32_TAIL": "1"}).split("RESULT")[1])
    for k in new:
        assert new[k]["e"] < FP32_TOL and new[k]["e2"] < FP32_TOL and old[k]["e"] < FP32_TOL, (k, new[k]["e"], new[k]["e2"], old[k]["e"])
        assert new[k]["e"] > 1e-12
        assert rel_err(np.array(new[k]["z"]), np.array(old[k]["z"]))[1] < FP32_TOL


def test_piecewise_staging_of_a_large_pinned_batch_on_the_gemm_chain(pkg):
    """GEMM-chain path, pinned source >= 256 MB: the observation matrix crosses PCIe in pieces with an event each and the first
    FVP's chunk loop waits per piece. With chunks no larger than a piece the chunking is the resident one and the result is
    bitwise the resident batch's, on the first and later FVPs; with the automatic (larger) chunk the streamed FVP walks the batch
    piece by piece -- same sums grouped differently, equal to rounding -- and every later FVP is bitwise the resident one again."""
    import torch
    layers, ac = [376, 64, 64, 17], "lttl"
    n = 200_000
    theta = pkg.synth.make_model(layers, 8)
    rng = np.random.default_rng(8)
    obs = rng.standard_normal((n, layers[0]))
    std = np.exp(theta[-layers[-1]:])
    v = rng.uniform(0, 1, theta.size)
    pinned = torch.from_numpy(obs).pin_memory()
    with pkg.Context(layers, ac) as ctx:
        ctx.set_model(theta)
        ctx.set_chunk(50_000)                                  # 4 chunks, pieces of ~89 k rows: chunks straddle pieces
        ctx.set_batch(obs, std)                                # pageable: plain synchronous copy
        z_ref = ctx.fvp(v, 0.1)
        for _ in range(2):
            ctx.set_batch(pinned.numpy(), std)                 # pinned: piecewise, the FVP below starts before the copy ends
            z1 = ctx.fvp(v, 0.1)
            z2 = ctx.fvp(v, 0.1)
            assert np.array_equal(z1, z_ref) and np.array_equal(z2, z_ref)
        ctx.set_batch(pinned.numpy(), std)
        x, info = ctx.cg(0.01 * v, 3, 0.0, 0.1)
        ctx.set_batch(obs, std)
        x_ref, _ = ctx.cg(0.01 * v, 3, 0.0, 0.1)
        assert info.cg_iters == 3 and np.array_equal(x, x_ref)
        ctx.set_chunk(0)                                       # automatic chunk (the whole batch): the streamed FVP steps by piece
        ctx.set_batch(obs, std)
        z_ref = ctx.fvp(v, 0.1)
        ctx.set_batch(pinned.numpy(), std)
        z1 = ctx.fvp(v, 0.1)
        z2 = ctx.fvp(v, 0.1)
        assert rel_err(z1, z_ref)[0] < 1e-13 and np.array_equal(z2, z_ref)


@pytest.mark.parametrize("layers,ac,n", [([376, 256, 256, 17], "lttl", 1500),      # Humanoid width: ragged 128-row tiles, 376 = 23.5 boxes
                                          ([40, 64, 48, 6], "lttl", 2100),          # widths that are not multiples of the 64-column tile
                                          ([18, 34, 22, 4], "ltsl", 900),           # boxes cut by the matrix edge in both directions
                                          ([17, 64, 64, 6], "lttl", 1300)])         # odd input width: layer 0 stays on the cp.async kernel
def test_tma_fed_chain_kernels_match_the_cp_async_ones(pkg, oracle, layers, ac, n):
    """gemm_chain_tma.cu: the outer-product and backward GEMMs fetch their tiles by TMA into swizzled boxes and walk the
    contraction index in a permuted order. Same sums as the cp.async kernels (TRPO_NO_CHAIN_TMA=1), grouped differently inside a
    32-wide k-step: equal to rounding, and both within the FP64 tolerance of the oracle; several chunks accumulate in place."""
    import os
    seed = 900 + sum(layers)
    theta = pkg.synth.make_model(layers, seed)
    batch = pkg.synth.make_batch(layers, ac, theta, n, seed)
    batch["Mean"] = oracle.forward(layers, ac, theta, batch["Observ"])
    vec = pkg.synth.make_vectors(layers, seed)
    z_ref = oracle.fvp(layers, ac, theta, batch["Std"], batch["Observ"], 0.1, vec["v"])
    pg_ref = oracle.policy_gradient(layers, ac, theta, batch["Observ"], batch["Mean"], batch["Action"], batch["Advantage"])
    out = {}
    try:
        for tag in ("tma", "cp_async"):
            if tag == "cp_async":
                os.environ["TRPO_NO_CHAIN_TMA"] = "1"
            for chunk in (0, 512):
                with pkg.Context(layers, ac) as ctx:
                    ctx.set_path(pkg.api.PATH_GEMM_CHAIN)
                    ctx.set_chunk(chunk)
                    ctx.set_model(theta)
                    ctx.set_batch(batch["Observ"], batch["Std"], batch["Mean"], batch["Action"], batch["Advantage"])
                    out[tag, chunk] = (ctx.fvp(vec["v"], 0.1), ctx.policy_gradient())
    finally:
        os.environ.pop("TRPO_NO_CHAIN_TMA", None)
    for key, (z, pg) in out.items():
        assert rel_err(z, z_ref)[0] < FVP_TOL and rel_err(z, z_ref)[1] < FVP_TOL, (key, rel_err(z, z_ref))
        assert rel_err(pg, pg_ref)[0] < FVP_TOL, (key, rel_err(pg, pg_ref))
    for chunk in (0, 512):
        assert rel_err(out["tma", chunk][0], out["cp_async", chunk][0])[0] < 1e-13
        assert rel_err(out["tma", chunk][1], out["cp_async", chunk][1])[0] < 1e-13
