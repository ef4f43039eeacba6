"""GPU parity tests for the steps either side of the update: rollout staging,
return / GAE advantage, the baseline objective as a libLBFGS callback, the binary batch file, and the whole training-loop
body against what the unmodified reference's TRPO_Lightweight wrote to its result files (tests/golden/lightweight.npz)."""
import os

import numpy as np
import pytest

import lightweight_loop as lw
from conftest import GOLDEN, load_synth, rel_err
from oracle_lib import Reference

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLDEN, "lightweight.npz")))


def _padded(x, n=lw.PADDED):
    out = np.zeros(n)
    out[:x.size] = x
    return out


@pytest.mark.parametrize("chain", [False, True])
def test_vf_evaluate_matches_the_reference_callback(pkg, gold, chain):
    """trpo_vf_evaluate against the output of the reference's own `evaluate` (TRPO_Baseline.c:29) on the first batch."""
    x = _padded(gold["x_base0"])
    with pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC) as ctx:
        if chain:
            ctx.set_path(pkg.api.PATH_GEMM_CHAIN)
        ctx.set_batch(gold["it0_Observ"], gold["it0_Std"])
        with pkg.ValueFunction(ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC) as vf:
            assert vf.num_params == 561
            vf.bind_batch(lw.EP_LEN)
            pred = vf.predict(x, lw.NUM_EP * lw.EP_LEN)
            assert rel_err(pred, gold["it0_ref_evaluate_predict"])[0] < TOL
            vf.set_target(gold["it0_Return"])
            fx, g = vf.evaluate(x)
            assert abs(fx - float(gold["it0_ref_evaluate_fx"])) < TOL * abs(float(gold["it0_ref_evaluate_fx"]))
            e_max, e_l2 = rel_err(g, gold["it0_ref_evaluate_g"])
            assert e_max < TOL and e_l2 < TOL, (e_max, e_l2)
            assert np.all(g[561:] == 0)
            # deterministic: the same call twice gives the same bits
            fx2, g2 = vf.evaluate(x)
            assert fx2 == fx and np.array_equal(g, g2)


def test_vf_evaluate_needs_a_target_and_a_matching_network(pkg, gold):
    with pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC) as ctx:
        ctx.set_batch(gold["it0_Observ"], gold["it0_Std"])
        with pytest.raises(RuntimeError):
            pkg.ValueFunction(ctx, [15, 16, 16, 1], "lttl")            # input must be ObservSpaceDim + 1
        with pytest.raises(RuntimeError):
            pkg.ValueFunction(ctx, [16, 16, 16, 2], "lttl")            # one output
        with pytest.raises(RuntimeError):
            pkg.ValueFunction(ctx, [16, 16, 16, 1], "lsst")            # TRPO_Baseline.c knows 'l' and 't' only
        with pkg.ValueFunction(ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC) as vf:
            with pytest.raises(RuntimeError):
                vf.bind_batch(7)                                        # 3000 is not a multiple of 7
            vf.bind_batch(lw.EP_LEN)
            with pytest.raises(RuntimeError):
                vf.evaluate(_padded(gold["x_base0"]))                   # no target yet


@pytest.mark.parametrize("vf_layers,ac,num_ep,ep_len", [([5, 7, 1], "ltl", 3, 11), ([9, 12, 6, 1], "lttl", 5, 8),
                                                        ([4, 6, 5, 3, 1], "ltltl", 2, 30), ([18, 64, 64, 1], "lttl", 7, 33),
                                                        ([3, 130, 70, 1], "lttl", 40, 1)])
def test_vf_evaluate_other_shapes_against_oracle(pkg, oracle, vf_layers, ac, num_ep, ep_len):
    rng = np.random.default_rng(17)
    N, O = num_ep * ep_len, vf_layers[0] - 1
    npar = sum(vf_layers[i] * vf_layers[i + 1] + vf_layers[i + 1] for i in range(len(vf_layers) - 1))
    x = _padded(rng.normal(size=npar) * 0.4, (npar + 15) // 16 * 16)
    obs = rng.normal(size=(N, O))
    tgt = rng.normal(size=N) * 3
    f_ref, g_ref, p_ref = oracle.vf_evaluate(vf_layers, ac, x, obs, tgt, num_ep, ep_len)
    pol_layers = [O, 4, 2]
    with pkg.Context(pol_layers, "ltl") as ctx:
        ctx.set_batch(obs, np.ones(2))
        with pkg.ValueFunction(ctx, vf_layers, ac) as vf:
            vf.bind_batch(ep_len)
            vf.set_target(tgt)
            fx, g = vf.evaluate(x)
            pred = vf.predict(x, N)
    assert rel_err(pred, p_ref)[0] < TOL
    assert abs(fx - f_ref) < TOL * abs(f_ref)
    assert rel_err(g, g_ref)[0] < TOL


def test_advantage_matches_oracle_on_the_reference_batch(pkg, gold):
    """Return / GAE / standardisation (TRPO_Lightweight.c:565-653) on the first rollout batch of the reference's run."""
    N = lw.NUM_EP * lw.EP_LEN
    with pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC) as ctx:
        ctx.set_model(gold["theta0"])
        ctx.set_rollout(lw.NUM_EP, lw.EP_LEN, gold["it0_Observ"], gold["it0_Std"], gold["it0_Mean"], gold["it0_Action"],
                        gold["it0_Reward"])
        with pkg.ValueFunction(ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC) as vf:
            ret, adv = vf.advantage(_padded(gold["x_base0"]), N, lw.GAMMA, lw.LAM)
            assert rel_err(ret, gold["it0_Return"])[0] < TOL
            assert rel_err(adv, gold["it0_Advantage"])[0] < TOL
            # the return is now the baseline's regression target, the advantage the policy batch's
            fx, g = vf.evaluate(_padded(gold["x_base0"]))
            assert abs(fx - float(gold["it0_ref_evaluate_fx"])) < TOL * abs(float(gold["it0_ref_evaluate_fx"]))
            assert rel_err(g, gold["it0_ref_evaluate_g"])[0] < TOL
        # ... so the TRPO update needs nothing else from the host: it lands on the reference loop's first update
        theta1, info = ctx.update(0.1)
    assert info.ls_accepted == 1
    assert np.abs(theta1 - gold["it0_theta"]).max() < 1e-9
    assert np.abs(theta1 - gold["ref_theta_iter1"]).max() < 1e-9       # what TRPO_Lightweight itself wrote (%.14f)


@pytest.mark.parametrize("num_ep,ep_len", [(1, 1), (3, 31), (5, 32), (4, 33), (2, 64), (9, 257), (1000, 1000)])
def test_gae_episode_lengths(pkg, oracle, num_ep, ep_len):
    """Chunked scan of the return / advantage recurrences for episode lengths around the 32-step chunk size, and at
    BASELINE size (1 M steps) against the vectorised recurrence on the host."""
    rng = np.random.default_rng(ep_len)
    N, O = num_ep * ep_len, 3
    vf_layers, ac = [O + 1, 8, 1], "ltl"
    x = rng.normal(size=(4 * 8 + 8 + 8 + 1)) * 0.5
    obs = rng.normal(size=(N, O))
    reward = rng.normal(size=N) * 2 - 1
    with pkg.Context([O, 4, 2], "ltl") as ctx:
        ctx.set_rollout(num_ep, ep_len, obs, np.ones(2), np.zeros((N, 2)), np.zeros((N, 2)), reward)
        with pkg.ValueFunction(ctx, vf_layers, ac) as vf:
            ret, adv = vf.advantage(x, N, lw.GAMMA, lw.LAM)
            base = vf.predict(x, N)
    if N <= 4096:
        b_ref = oracle.vf_predict(vf_layers, ac, x, obs, num_ep, ep_len)
        assert rel_err(base, b_ref)[0] < TOL
        if N > 1:
            r_ref, a_ref = oracle.gae(reward, b_ref, num_ep, ep_len, lw.GAMMA, lw.LAM)
            assert rel_err(ret, r_ref)[0] < TOL and rel_err(adv, a_ref)[0] < TOL
        return
    R, V = reward.reshape(num_ep, ep_len), base.reshape(num_ep, ep_len)
    ret2, adv2 = np.zeros_like(R), np.zeros_like(R)
    acc_r, acc_a, nxt = np.zeros(num_ep), np.zeros(num_ep), np.zeros(num_ep)
    for t in range(ep_len - 1, -1, -1):
        acc_r = R[:, t] + lw.GAMMA * acc_r
        acc_a = (R[:, t] + lw.GAMMA * nxt - V[:, t]) + lw.GAMMA * lw.LAM * acc_a
        ret2[:, t], adv2[:, t], nxt = acc_r, acc_a, V[:, t]
    adv2 = adv2.reshape(N)
    adv2 = (adv2 - adv2.mean()) / adv2.std()
    assert rel_err(ret, ret2.reshape(N))[0] < TOL and rel_err(adv, adv2)[0] < TOL
    assert abs(adv.mean()) < 1e-12 and abs(adv.std() - 1) < 1e-12


class GpuBackend:
    """The compute steps of lightweight_loop.run on the GPU library; libLBFGS calls trpo_vf_evaluate natively."""

    def __init__(self, pkg, chain=False):
        self.ctx = pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC)
        if chain:
            self.ctx.set_path(pkg.api.PATH_GEMM_CHAIN)
        self.vf = pkg.ValueFunction(self.ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC)

    def advantage(self, batch, x_base):
        self.ctx.set_rollout(lw.NUM_EP, lw.EP_LEN, batch["Observ"], batch["Std"], batch["Mean"], batch["Action"], batch["Reward"])
        return self.vf.advantage(x_base, lw.NUM_EP * lw.EP_LEN, lw.GAMMA, lw.LAM)

    def vf_callback(self, batch, target):
        return self.vf.callback_pointer()            # the return is already installed as the target

    def update(self, theta, batch, adv, damping):
        self.ctx.set_model(theta)
        out, _ = self.ctx.update(damping)            # the standardised advantage is already in the device batch
        return out

    def close(self):
        self.vf.close()
        self.ctx.close()


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref (the reference's vendored libLBFGS) not built")
@pytest.mark.parametrize("chain", [False, True])
def test_training_loop_body_matches_the_reference_result_files(pkg, oracle, gold, chain):
    """Three iterations of rollouts -> advantage -> baseline fit -> TRPO update with every compute step on the GPU (the
    reference's libLBFGS drives trpo_vf_evaluate as its callback) land on the parameters the unmodified
    TRPO_Lightweight wrote to its result files (%.14f)."""
    ref = Reference()
    be = GpuBackend(pkg, chain)
    trace = []
    try:
        lw.run(be, oracle, ref, gold["theta0"], gold["x_base0"], 3, trace=trace)
    finally:
        be.close()
    assert rel_err(trace[0]["ret"], gold["it0_Return"])[0] < TOL
    assert rel_err(trace[0]["adv"], gold["it0_Advantage"])[0] < TOL
    assert rel_err(trace[0]["x_base"], gold["it0_x_base_fitted"])[0] < 1e-8     # 25 L-BFGS iterations amplify rounding
    for i in (1, 2, 3):
        e = np.abs(trace[i - 1]["theta"] - gold[f"ref_theta_iter{i}"]).max()
        assert e < 1e-8, (i, e)


def test_rollout_producer_matches_the_reference_batch(pkg, gold):
    """trpo_ctx_rollout_arm fed with the rand() stream of srand(0) reproduces the first batch of the reference's run
    (tests/golden/lightweight.npz holds the oracle's copy, pinned by the result files)."""
    import ctypes as C
    C.CDLL(None).srand(0)
    draws = lw.libc_rand_draws(lw.NUM_EP * (3 + 6 * lw.EP_LEN))
    N = lw.NUM_EP * lw.EP_LEN
    with pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC) as ctx:
        ctx.set_model(gold["theta0"])
        ctx.rollout_arm(lw.NUM_EP, lw.EP_LEN, draws)
        b = ctx.get_rollout(N)
        for k in ("Observ", "Mean", "Action", "Reward"):
            err = np.abs(b[k] - gold["it0_" + k]).max() / np.abs(gold["it0_" + k]).max()
            assert err < TOL, (k, err)
        # the produced batch is staged: advantage and update run on it without any host copy of the observations
        with pkg.ValueFunction(ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC) as vf:
            ret, adv = vf.advantage(_padded(gold["x_base0"]), N, lw.GAMMA, lw.LAM)
        assert rel_err(ret, gold["it0_Return"])[0] < 1e-9 and rel_err(adv, gold["it0_Advantage"])[0] < 1e-9
        theta1, info = ctx.update(0.1)
    assert np.abs(theta1 - gold["ref_theta_iter1"]).max() < 1e-8


def test_rollout_producer_device_generator(pkg, gold):
    """Counter-based generator: deterministic per seed, different across seeds, and the sampled actions are
    Mean + Std * N(0,1) with the object inside the simulator's box (TRPO_Lightweight.c:391-393)."""
    num_ep, ep_len = 2000, 50
    N = num_ep * ep_len
    with pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC) as ctx:
        ctx.set_model(gold["theta0"])
        ctx.rollout_arm(num_ep, ep_len, None, seed=7)
        a = ctx.get_rollout(N)
        ctx.rollout_arm(num_ep, ep_len, None, seed=7)
        a2 = ctx.get_rollout(N)
        ctx.rollout_arm(num_ep, ep_len, None, seed=8)
        b = ctx.get_rollout(N)
    for k in a:
        assert np.array_equal(a[k], a2[k])
    assert not np.array_equal(a["Action"], b["Action"])
    std = np.exp(gold["theta0"][-3:])
    z = (a["Action"] - a["Mean"]) / std
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    obj = a["Observ"][:, 12:15]
    assert obj[:, 0].min() >= 0.084 and obj[:, 0].max() <= 0.16 and obj[:, 1].min() >= -0.05 and obj[:, 1].max() <= 0.05
    assert obj[:, 2].min() >= 0 and obj[:, 2].max() <= 0.1
    assert np.all(a["Reward"] < 0) and np.isfinite(a["Reward"]).all()
    # the first step of every episode starts from the reset pose (TRPO_Lightweight.c:357-385)
    first = a["Observ"][::ep_len]
    assert np.allclose(first[:, :12], [0, 0, 0.01768, 0, 0, 0.07518, 0.07375, 0, 0.07518, 0.11315, 0, 0.06268])


class GpuProducerBackend(GpuBackend):
    """Rollouts produced on the device too: only the rand() draws and the P-length vectors cross PCIe."""

    def rollout(self, theta):
        self.ctx.set_model(theta)
        self.ctx.rollout_arm(lw.NUM_EP, lw.EP_LEN, lw.libc_rand_draws(lw.NUM_EP * (3 + 6 * lw.EP_LEN)))
        b = self.ctx.get_rollout(lw.NUM_EP * lw.EP_LEN)         # for the trace only
        b["Std"] = np.exp(theta[-3:])
        return b

    def advantage(self, batch, x_base):
        return self.vf.advantage(x_base, lw.NUM_EP * lw.EP_LEN, lw.GAMMA, lw.LAM)


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref (the reference's vendored libLBFGS) not built")
def test_training_loop_with_device_rollouts_matches_the_reference_result_files(pkg, oracle, gold):
    """The whole loop of TRPO_Lightweight -- simulator, advantage, baseline fit, TRPO update -- with every compute step on
    the GPU, fed with the rand() stream the reference consumes: same parameters after 3 iterations."""
    ref = Reference()
    be = GpuProducerBackend(pkg)
    trace = []
    try:
        lw.run(be, oracle, ref, gold["theta0"], gold["x_base0"], 3, trace=trace)
    finally:
        be.close()
    assert np.abs(trace[0]["batch"]["Observ"] - gold["it0_Observ"]).max() < 1e-10
    for i in (1, 2, 3):
        e = np.abs(trace[i - 1]["theta"] - gold[f"ref_theta_iter{i}"]).max()
        assert e < 1e-7, (i, e)


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref (the reference's TRPO_Lightweight and libLBFGS) not built")
def test_trpo_lightweight_gpu_entry_point_writes_the_reference_result_files(pkg, gold, tmp_path, monkeypatch, capfd):
    """TRPO_Lightweight_GPU (drop-in for TRPO_Lightweight / TRPO_Lightweight_FPGA, TRPO.h:110,113) against the unmodified
    reference run here from the same model / baseline files: same result files, same log lines."""
    import oracle_lib
    mf, bf = str(tmp_path / "model.txt"), str(tmp_path / "base.txt")
    pkg.textio.write_model(mf, gold["theta0"])
    pkg.textio.write_model(bf, gold["x_base0"])
    monkeypatch.setenv("TRPO_LBFGS_LIB", os.path.join(oracle_lib.ORACLE_DIR, "_ref", "libtrpo_ref.so"))
    monkeypatch.chdir(tmp_path)                     # the reference's result-file name buffer is 30 bytes: keep it short
    t_ref = Reference().lightweight(mf, bf, "ref", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, 3)
    out_ref = capfd.readouterr().out
    t_gpu = pkg.api.TRPO_Lightweight_GPU(mf, bf, "gpu", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, 3)
    out_gpu = capfd.readouterr().out
    assert t_ref > 0 and t_gpu > 0
    ref = np.loadtxt(tmp_path / "ref002.txt")
    got = np.loadtxt(tmp_path / "gpu002.txt")
    assert np.abs(ref - gold["ref_theta_iter3"]).max() == 0
    assert np.abs(got - ref).max() < 1e-7
    assert os.path.exists(tmp_path / "gpu000.txt") and not os.path.exists(tmp_path / "gpu001.txt")   # iter % 100 == 0, last
    # same log: three reward lines with the same statistics, same number of CG lines, same line-search outcomes
    rew_ref = [l for l in out_ref.splitlines() if l.startswith("[INFO] Iteration")]
    rew_gpu = [l for l in out_gpu.splitlines() if l.startswith("[INFO] Iteration")]
    assert rew_ref == rew_gpu and len(rew_gpu) == 3
    assert out_ref.count("CG Iter[") == out_gpu.count("CG Iter[")
    assert out_ref.count("a/e/r") == out_gpu.count("a/e/r")
    # a missing baseline file fails like the reference: message + -1
    t = pkg.api.TRPO_Lightweight_GPU(mf, str(tmp_path / "nope.txt"), "gpu", lw.ARM_LAYERS, lw.ARM_ACFUNC, 0.1, 1)
    assert t == -1.0 and "[ERROR] Cannot open BaselineFile" in capfd.readouterr().err


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref (the reference's libLBFGS) not built")
def test_c_harness_lightweight_mode(pkg, gold, tmp_path):
    """host/trpo_test_main.c in the shape of Test_TRPO_Lightweight_FPGA (TRPOCpuCode.c:433-465): a plain C program, the
    C-ABI library and the reference's libLBFGS -- no Python in the loop."""
    import subprocess
    import oracle_lib
    exe = os.path.join(os.path.dirname(pkg.api.library_path()), "trpo_test_gpu")
    mf, bf = str(tmp_path / "model.txt"), str(tmp_path / "base.txt")
    pkg.textio.write_model(mf, gold["theta0"])
    pkg.textio.write_model(bf, gold["x_base0"])
    env = dict(os.environ, TRPO_LBFGS_LIB=os.path.join(oracle_lib.ORACLE_DIR, "_ref", "libtrpo_ref.so"))
    out = subprocess.run([exe, "lightweight", mf, bf, str(tmp_path / "res"), "2"], capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    assert "[INFO] Iteration 1, Episode Rewards Mean = -476.223178, Std = 55.417171" in out.stdout   # the reference's own log
    got = np.loadtxt(tmp_path / "res001.txt")
    assert np.abs(got - gold["ref_theta_iter2"]).max() < 1e-7
    # without a libLBFGS in reach the entry point says so and fails
    env.pop("TRPO_LBFGS_LIB")
    out = subprocess.run([exe, "lightweight", mf, bf, str(tmp_path / "res"), "1"], capture_output=True, text=True,
                         timeout=300, env=env)
    assert out.returncode != 0 and "libLBFGS not found" in out.stderr


def test_rollout_api_argument_errors(pkg, gold, tmp_path):
    """Bad arguments fail loudly (RuntimeError carrying trpo_last_error) and leave the context usable."""
    with pkg.Context([17, 64, 64, 6], "lttl") as ctx:
        ctx.set_model(pkg.synth.make_model([17, 64, 64, 6], 3))
        with pytest.raises(RuntimeError, match="15-...-3"):
            ctx.rollout_arm(4, 10, None, 1)                 # the simulator's observation / action layout
    with pkg.Context(lw.ARM_LAYERS, lw.ARM_ACFUNC) as ctx:
        ctx.set_model(gold["theta0"])
        with pytest.raises(RuntimeError):
            ctx.rollout_arm(0, 10, None, 1)
        with pytest.raises(RuntimeError):
            ctx.get_rollout(10)                             # nothing staged yet
        with pytest.raises(RuntimeError):
            ctx.set_batch_file(str(tmp_path / "missing.bin"))
        with pkg.ValueFunction(ctx, lw.ARM_VF_LAYERS, lw.ARM_ACFUNC) as vf:
            with pytest.raises(RuntimeError, match="set_rollout"):
                vf.advantage(_padded(gold["x_base0"]), 10, lw.GAMMA, lw.LAM)      # no rollout staged
            with pytest.raises(RuntimeError):
                vf.predict(_padded(gold["x_base0"]), 10)                          # not bound to a batch
        ctx.rollout_arm(2, 5, None, 3)                      # still works afterwards
        assert np.isfinite(ctx.get_rollout(10)["Reward"]).all()


def test_binary_batch_file_staging(pkg, tmp_path):
    """trpo_ctx_set_batch_file and a binary DataFile behind the file-based entry points give the same bits as the text
    file / host arrays."""
    s = load_synth("mlp64")
    layers, ac = s["layers"], s["acfunc"]
    N = s["Observ"].shape[0]
    bf, df, mf = str(tmp_path / "d.bin"), str(tmp_path / "d.txt"), str(tmp_path / "m.txt")
    pkg.api.batch_file_write(bf, s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
    pkg.textio.write_data(df, s["Mean"], s["Std"], s["Observ"], s["Action"], s["Advantage"])
    pkg.textio.write_model(mf, s["theta"])
    with pkg.Context(layers, ac) as ctx:
        ctx.set_model(s["theta"])
        ctx.set_batch(s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
        z0, b0 = ctx.fvp(s["v"], 0.1), ctx.policy_gradient()
        ctx.set_batch_file(bf)
        z1, b1 = ctx.fvp(s["v"], 0.1), ctx.policy_gradient()
        assert np.array_equal(z0, z1) and np.array_equal(b0, b1)
        ctx.set_batch_file(bf, 100)
        z2 = ctx.fvp(s["v"], 0.1)
        ctx.set_batch(s["Observ"][:100], s["Std"])
        assert np.array_equal(z2, ctx.fvp(s["v"], 0.1))
        with pytest.raises(RuntimeError):
            ctx.set_batch_file(df)
    x_txt, t_txt = pkg.CG_GPU(mf, df, layers, ac, N, 0.1, s["b"])
    x_bin, t_bin = pkg.CG_GPU(mf, bf, layers, ac, N, 0.1, s["b"])
    assert t_txt >= 0 and t_bin >= 0 and np.array_equal(x_txt, x_bin)
    assert rel_err(x_bin, s["ref_cg"])[0] < 1e-8
    u_bin, t = pkg.TRPO_Update_GPU(mf, bf, layers, ac, N, 0.1)
    assert t >= 0 and rel_err(u_bin, s["ref_update"])[0] < 1e-8
