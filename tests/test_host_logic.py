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
