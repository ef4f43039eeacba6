"""CPU tests for rows f-3 / f-4: the oracle's rollout / GAE / baseline-objective restatements against the committed
golden outputs of the compiled reference (tests/golden/lightweight.npz) and against oracle/_ref when it is present."""
import os

import numpy as np
import pytest

import lightweight_loop as lw
from oracle_lib import Reference

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lightweight.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


def _padded(x):
    out = np.zeros(lw.PADDED)
    out[:x.size] = x
    return out


def test_vf_evaluate_matches_reference_evaluate_bit_exact(oracle, gold):
    """oracle_vf_evaluate == the reference's libLBFGS callback (TRPO_Baseline.c:29) on the first rollout batch."""
    fx, g, pred = oracle.vf_evaluate(lw.ARM_VF_LAYERS, lw.ARM_ACFUNC, _padded(gold["x_base0"]), gold["it0_Observ"],
                                     gold["it0_Return"], lw.NUM_EP, lw.EP_LEN)
    assert fx == float(gold["it0_ref_evaluate_fx"])
    assert np.array_equal(g, gold["it0_ref_evaluate_g"])
    assert np.array_equal(pred, gold["it0_ref_evaluate_predict"])
    assert np.all(g[561:] == 0)


def test_gae_matches_closed_form(oracle, gold):
    """Return and GAE are discounted suffix sums; the reference's pow() form and the recurrence agree to rounding."""
    r, N = gold["it0_Reward"], lw.NUM_EP * lw.EP_LEN
    base = oracle.vf_predict(lw.ARM_VF_LAYERS, lw.ARM_ACFUNC, _padded(gold["x_base0"]), gold["it0_Observ"], lw.NUM_EP, lw.EP_LEN)
    assert np.array_equal(base, gold["it0_ref_evaluate_predict"])
    ret, adv = oracle.gae(r, base, lw.NUM_EP, lw.EP_LEN, lw.GAMMA, lw.LAM)
    assert np.array_equal(ret, gold["it0_Return"]) and np.array_equal(adv, gold["it0_Advantage"])
    R, V = r.reshape(lw.NUM_EP, lw.EP_LEN), base.reshape(lw.NUM_EP, lw.EP_LEN)
    ret2, adv2 = np.zeros_like(R), np.zeros_like(R)
    nxt_v = np.zeros(lw.NUM_EP)
    acc_r, acc_a = np.zeros(lw.NUM_EP), np.zeros(lw.NUM_EP)
    for t in range(lw.EP_LEN - 1, -1, -1):
        acc_r = R[:, t] + lw.GAMMA * acc_r
        delta = R[:, t] + lw.GAMMA * nxt_v - V[:, t]
        acc_a = delta + lw.GAMMA * lw.LAM * acc_a
        ret2[:, t], adv2[:, t], nxt_v = acc_r, acc_a, V[:, t]
    adv2 = adv2.reshape(N)
    adv2 = (adv2 - adv2.mean()) / adv2.std()
    assert np.abs(ret2.reshape(N) - ret).max() < 1e-11 * np.abs(ret).max()
    assert np.abs(adv2 - adv).max() < 1e-11
    assert abs(adv.mean()) < 1e-13 and abs(adv.std() - 1) < 1e-13


def test_reward_stats(oracle, gold):
    m, s = oracle.reward_stats(gold["it0_Reward"], lw.NUM_EP, lw.EP_LEN)
    ep = gold["it0_Reward"].reshape(lw.NUM_EP, lw.EP_LEN).sum(axis=1)
    assert abs(m - ep.mean()) < 1e-10 and abs(s - ep.std()) < 1e-10
    assert abs(m - (-512.069339)) < 1e-6 and abs(s - 43.393057) < 1e-6      # the reference's own log line, iteration 0


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_loop_lands_on_the_reference_result_files(oracle, gold):
    """Rollout + GAE + baseline fit (reference libLBFGS, oracle objective) + update, three iterations: the parameters
    equal what the unmodified TRPO_Lightweight wrote to its result files (printed with %.14f)."""
    ref = Reference()
    trace = []
    got = lw.run(lw.OracleBackend(oracle), oracle, ref, gold["theta0"], gold["x_base0"], 3, trace=trace)
    for i in (1, 2, 3):
        assert np.abs(trace[i - 1]["theta"] - gold[f"ref_theta_iter{i}"]).max() < 1e-14 + 5e-15
    assert np.array_equal(trace[0]["batch"]["Observ"], gold["it0_Observ"])
    assert np.array_equal(trace[0]["batch"]["Action"], gold["it0_Action"])
    assert np.array_equal(got, gold["it2_theta"])


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_vf_evaluate_vs_live_reference_other_shapes(oracle):
    ref = Reference()
    rng = np.random.default_rng(5)
    for vf_layers, ac, ne, el in (([5, 7, 1], "ltl", 3, 11), ([9, 12, 6, 1], "lttl", 5, 8), ([4, 6, 5, 3, 1], "ltltl", 2, 30)):
        npar = sum(vf_layers[i] * vf_layers[i + 1] + vf_layers[i + 1] for i in range(len(vf_layers) - 1))
        n = (npar + 15) // 16 * 16
        x = np.zeros(n)
        x[:npar] = rng.normal(size=npar) * 0.4
        obs = rng.normal(size=(ne * el, vf_layers[0] - 1))
        tgt = rng.normal(size=ne * el) * 3
        f1, g1, p1 = oracle.vf_evaluate(vf_layers, ac, x, obs, tgt, ne, el)
        f2, g2, p2 = ref.vf_evaluate(vf_layers, ac, x, obs, tgt, ne, el)
        assert f1 == f2 and np.array_equal(g1, g2) and np.array_equal(p1, p2)
