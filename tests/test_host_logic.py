"""CPU tests of the host-side logic: reference text formats, synthetic generator, flat layout, sample sharding."""
import numpy as np

from conftest import load_synth


def test_text_formats_round_trip(pkg, oracle, tmp_path):
    s = load_synth("net3")
    mf, df = str(tmp_path / "m.txt"), str(tmp_path / "d.txt")
    pkg.textio.write_model(mf, s["theta"])
    pkg.textio.write_data(df, s["Mean"], s["Std"], s["Observ"], s["Action"], s["Advantage"])
    assert np.array_equal(pkg.textio.read_model(mf, s["theta"].size), s["theta"])
    d = pkg.textio.read_data(df, s["layers"], s["Observ"].shape[0])
    for k in ("Mean", "Std", "Observ", "Action", "Advantage"):
        assert np.array_equal(d[k], s[k]), k
    # and the C loader of the oracle (same fscanf format as the reference) parses the same numbers
    assert np.array_equal(oracle.load_model(mf, s["layers"], s["acfunc"]), s["theta"])


def test_flat_layout_is_augmented_matrices(pkg):
    layers = [3, 4, 2]
    w, b, ls = pkg.synth.layout(layers)
    assert w == [0, 16] and b == [12, 24] and ls == 26
    assert pkg.synth.num_params(layers) == 28


def test_synth_forward_matches_oracle(pkg, oracle):
    layers, ac = [6, 8, 7, 5, 2], "lstso"
    theta = pkg.synth.make_model(layers, 7)
    batch = pkg.synth.make_batch(layers, ac, theta, 50, 7)
    ref = oracle.forward(layers, ac, theta, batch["Observ"])
    assert np.allclose(batch["Mean"], ref, rtol=1e-12, atol=1e-14)
    assert np.allclose(batch["Std"], np.exp(theta[-2:]))
    assert abs(batch["Advantage"].mean()) < 1e-12 and abs(batch["Advantage"].std() - 1) < 1e-12


def test_shard_bounds_cover_all_samples(pkg):
    from bench import shard_bounds
    for n, g in ((10, 3), (1_000_000, 8), (7, 8), (3150, 2)):
        cuts = [shard_bounds(n, g, r) for r in range(g)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(g - 1))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1


def test_branch_free_tanh_algorithm_accuracy():
    """The fused kernel's tanh (csrc/fvp_fused.cu: tanh_vec) restated in numpy: |error| < 4e-16 against libm."""
    MAGIC, L2E2 = 6755399441055744.0, 2.8853900817779268
    HI, LO = 0.3465735901845619, 9.541074646352939e-11
    C = [1.0, 2.0, 2.0, 1.3333333333333333, 0.6666666666666666, 0.26666666666666666, 0.08888888888888889,
         0.025396825396825397, 0.006349206349206349, 0.0014109347442680777, 0.0002821869488536155,
         5.130671797338464e-05, 8.551119662230774e-06]
    x = np.concatenate([np.linspace(-25, 25, 100001), np.logspace(-14, 1.3, 5000), -np.logspace(-14, 1.3, 5000)])
    a = np.minimum(np.abs(x), 20.0)
    nf = (a * L2E2 + MAGIC) - MAGIC
    h = (a - nf * HI) - nf * LO
    q = np.full_like(a, C[12])
    for k in range(11, -1, -1):
        q = q * h + C[k]
    s = np.ldexp(q, nf.astype(int)) + 1.0
    y = (1.0 / s).astype(np.float32).astype(np.float64)          # stands in for MUFU.RCP64H (>= 20 good bits)
    for _ in range(2):
        y = y + y * (1.0 - s * y)
    got = np.copysign(1.0 - 2.0 * y, x)
    assert np.abs(got - np.tanh(x)).max() < 4e-16
