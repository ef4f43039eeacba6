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
    """The kernels' tanh (csrc/dmma_common.cuh: tanh_vec) restated in numpy: |error| < 4e-16 against libm."""
    MAGIC, INV = 6755399441055744.0, 92.33248261689366
    HI, LO = 0.010830424667801708, 2.8447437476627285e-11
    tab = np.exp2(np.arange(64) / 64.0)
    x = np.concatenate([np.linspace(-25, 25, 100001), np.logspace(-14, 1.3, 5000), -np.logspace(-14, 1.3, 5000)])
    z = 2 * np.minimum(np.abs(x), 20.0)
    kf = (z * INV + MAGIC) - MAGIC
    k = kf.astype(np.int64)
    r = (z - kf * HI) - kf * LO
    assert np.abs(r).max() < 0.0055
    q = np.full_like(z, 1.0 / 120.0)
    for c in (1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0):
        q = q * r + c
    s = np.ldexp(q * tab[k & 63], (k >> 6).astype(int)) + 1.0
    y = (1.0 / s).astype(np.float32).astype(np.float64)          # stands in for MUFU.RCP64H (>= 20 good bits)
    for _ in range(2):
        y = y + y * (1.0 - s * y)
    got = np.copysign(1.0 - 2.0 * y, x)
    assert np.abs(got - np.tanh(x)).max() < 4e-16


def test_binary_batch_file_round_trip_and_text_conversion(pkg, tmp_path):
    """The binary data file (include/trpo_b200.h, host/trpo_batch_file.c) holds exactly what the text file parses to."""
    s = load_synth("net3")
    N, O, A = s["Observ"].shape[0], s["Observ"].shape[1], s["Std"].size
    bf = str(tmp_path / "d.bin")
    pkg.api.batch_file_write(bf, s["Observ"], s["Std"], s["Mean"], s["Action"], s["Advantage"])
    import os
    assert os.path.getsize(bf) == 64 + 8 * (A + N * O + 2 * N * A + N)
    d = pkg.api.batch_file_read(bf, N, O, A)
    for k in ("Mean", "Std", "Observ", "Action", "Advantage"):
        assert np.array_equal(d[k], s[k]), k
    # a prefix of the rows, as TRPOparam.NumSamples < file rows does with the text file
    d = pkg.api.batch_file_read(bf, 17, O, A)
    assert np.array_equal(d["Observ"], s["Observ"][:17]) and np.array_equal(d["Advantage"], s["Advantage"][:17])
    assert np.array_equal(d["Action"], s["Action"][:17])
    # text -> binary conversion parses the reference's row format (TRPO_FVP.c:739-760)
    df, bf2 = str(tmp_path / "d.txt"), str(tmp_path / "d2.bin")
    pkg.textio.write_data(df, s["Mean"], s["Std"], s["Observ"], s["Action"], s["Advantage"])
    pkg.api.batch_file_from_text(df, bf2, N, O, A)
    assert open(bf, "rb").read() == open(bf2, "rb").read()
    # asking for more rows than the file holds fails like a short text file does
    import pytest
    with pytest.raises(RuntimeError):
        pkg.api.batch_file_read(bf, N + 1, O, A)
    with pytest.raises(RuntimeError):
        pkg.api.batch_file_read(df, N, O, A)          # a text file is not a binary batch file


def test_chunked_scan_of_the_return_and_advantage_recurrences(oracle):
    """The algorithm of k_gae (csrc/rollout_kernels.cu) emulated lane by lane: 32-step chunks walked backwards, a 5-step
    shuffle scan inside a chunk, the carry from the chunk behind weighted a^(32 - lane). It must agree with the
    reference's O(EpLen^2) pow() sums (oracle_gae) for episode lengths around the chunk size."""
    gamma, lam = 0.995, 0.98
    rng = np.random.default_rng(0)

    def scan_episode(d, a):
        n = len(d)
        out = np.zeros(n)
        carry = 0.0
        for t0 in range(((n - 1) // 32) * 32, -1, -32):
            x = np.zeros(32)
            m = min(32, n - t0)
            x[:m] = d[t0:t0 + m]
            p, off = a, 1
            while off < 32:
                y = np.concatenate([x[off:], np.zeros(off)])        # __shfl_down, lanes past the end masked out
                x = x + p * y
                p, off = p * p, off * 2
            x = x + a ** (32 - np.arange(32)) * carry
            out[t0:t0 + m] = x[:m]
            carry = x[0]
        return out

    for num_ep, ep_len in ((3, 1), (2, 31), (2, 32), (3, 33), (2, 150), (1, 257)):
        N = num_ep * ep_len
        reward, base = rng.normal(size=N) * 2 - 1, rng.normal(size=N)
        ret = np.zeros(N)
        adv = np.zeros(N)
        for ep in range(num_ep):
            r, v = reward[ep * ep_len:(ep + 1) * ep_len], base[ep * ep_len:(ep + 1) * ep_len]
            delta = r + gamma * np.concatenate([v[1:], [0.0]]) - v
            ret[ep * ep_len:(ep + 1) * ep_len] = scan_episode(r, gamma)
            adv[ep * ep_len:(ep + 1) * ep_len] = scan_episode(delta, gamma * lam)
        if N > 1:
            adv = (adv - adv.mean()) / adv.std()
            r_ref, a_ref = oracle.gae(reward, base, num_ep, ep_len, gamma, lam)
            assert np.abs(ret - r_ref).max() < 1e-12 * max(1.0, np.abs(r_ref).max())
            assert np.abs(adv - a_ref).max() < 1e-11


def _sw128(row, col):
    """Byte offset of FP64 element (row, col) inside a SWIZZLE_128B box of 16 columns (gemm_chain_tma.cu)."""
    return row * 128 + (((col >> 1) ^ (row & 7)) << 4) + (col & 1) * 8


def _conflict_free(offsets):
    """An LDS.64 is served per half warp: its 16 lanes must hit 16 distinct 8-byte words of the 128-byte bank window."""
    for half in (offsets[:16], offsets[16:]):
        assert len({(o % 128) // 8 for o in half}) == 16, sorted((o % 128) // 8 for o in half)


def test_tma_fragment_addressing_is_a_conflict_free_permutation_of_k():
    """The TMA-fed GEMM kernels read their m8n8k4 fragments through the 128-byte swizzle with the contraction index permuted
    (gemm_chain_tma.cu header). Restated here lane by lane: every k of a k-step is visited exactly once, both operands of a DMMA
    see the same k in the same lane, and no half warp has a shared-memory bank conflict."""
    lanes = [(l >> 2, l & 3) for l in range(32)]                      # (g, t)
    # operands whose box rows are k (outer product): lane t of step q reads row 8*(q/2) + 2t + (q&1), column 8*ib + g
    seen = set()
    for q in range(8):
        ks = [8 * (q >> 1) + 2 * t + (q & 1) for t in range(4)]
        seen.update(ks)
        for ib in range(2):
            _conflict_free([_sw128(8 * (q >> 1) + 2 * t + (q & 1), 8 * ib + g) for g, t in lanes])
    assert seen == set(range(32))
    # operands whose box columns are k (backward GEMM, forward activations): lane t of step c0 reads column 2c0 + 8(t/2) + (t&1)
    seen = set()
    for c0 in range(4):
        cols = [2 * c0 + 8 * (t >> 1) + (t & 1) for t in range(4)]
        seen.update(cols)
        for i in range(4):
            _conflict_free([_sw128(8 * i + g, 2 * c0 + 8 * (t >> 1) + (t & 1)) for g, t in lanes])
    assert seen == set(range(16))
    # forward layer: the weights' box rows are k, stored through the row permutation of k_permute_rows16 (bit 0 <-> 1, 2 <-> 3)
    perm = [((r & 1) << 1) | ((r & 2) >> 1) | ((r & 4) << 1) | ((r & 8) >> 1) for r in range(16)]
    assert sorted(perm) == list(range(16)) and all(perm[perm[r]] == r for r in range(16))
    for c0 in range(4):
        for t in range(4):
            k_act = 2 * c0 + 8 * (t >> 1) + (t & 1)                   # the k the activation fragment of lane t holds
            rho = 8 * (c0 >> 1) + 2 * t + (c0 & 1)                    # the stored weight row that lane reads
            assert perm[rho] == k_act                                 # ... holds the same k
        for j in range(2):
            _conflict_free([_sw128(8 * (c0 >> 1) + 2 * t + (c0 & 1), 8 * j + g) for g, t in lanes])
    # un-permuted, lanes t and t + 2 would read rows 8 apart: same swizzle phase, same banks
    bad = [_sw128(2 * 0 + 8 * (t >> 1) + (t & 1), g) for g, t in lanes]
    assert len({(o % 128) // 8 for o in bad[:16]}) < 16
