"""The stream definitions of include/smm_stream.h (CPU side): Philox known-answer vectors, accuracy of
the fp64 transform against a 120-bit evaluation of its own definition, distribution sanity."""
import mpmath as mp
import numpy as np
from scipy import stats


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for c, k, want in kat:
        assert tuple(int(x) for x in oracle.philox(c, k)) == want


def test_neglog_accuracy(oracle):
    mp.mp.prec = 120
    rng = np.random.default_rng(1)
    us = list(rng.random(1500)) + [2.0 ** -52, 0.5, 1 - 2.0 ** -52, 0.70710678118, 2.0 ** -30 * 1.999, 1 - 2.0 ** -20]
    worst = max(abs(mp.mpf(oracle.neglog01(float(u))) + mp.log(mp.mpf(float(u)))) / -mp.log(mp.mpf(float(u))) for u in us)
    assert worst < 1e-14


def test_normal_pair_matches_its_definition(oracle):
    """z0,z1 = sqrt(-2 ln u1) (cos,sin)((pi/4) g) under the drawn symmetry, to ~4 ulp"""
    mp.mp.prec = 120
    rng = np.random.default_rng(2)
    worst = mp.mpf(0)
    for _ in range(1500):
        x, y, z, w = [int(v) for v in rng.integers(0, 2 ** 32, 4)]
        z0, z1 = oracle.normal_from_words(x, y, z, w)
        A = ((x << 20) | (y >> 12)) | 1
        u1 = mp.mpf(2 ** 52 - A) / 2 ** 52
        g = mp.mpf(((z << 20) | (w >> 12)) & (2 ** 49 - 1)) / 2 ** 49
        rad = mp.sqrt(-2 * mp.log(u1))
        c, s = mp.cos(mp.pi / 4 * g), mp.sin(mp.pi / 4 * g)
        if (z >> 29) & 1:
            c, s = s, c
        if (z >> 30) & 1:
            c = -c
        if (z >> 31) & 1:
            s = -s
        worst = max(worst, abs(mp.mpf(z0) - rad * c), abs(mp.mpf(z1) - rad * s))
    assert worst < 4e-15


def test_extreme_words(oracle):
    for words in [(0, 0, 0, 0), (0xffffffff,) * 4, (0xffffffff, 0xffffffff, 0, 0), (0, 0, 0xffffffff, 0xffffffff)]:
        z = oracle.normal_from_words(*words)
        assert np.all(np.isfinite(z)) and np.all(np.abs(z) < 8.6)


def test_normals_are_standard_normal(oracle):
    z = oracle.normals(1234, 0, 0, 1 << 28, 500_000)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5e-3           # the pair is uncorrelated
    assert stats.kstest(z[0::2] ** 2 + z[1::2] ** 2, "chi2", args=(2,)).pvalue > 1e-3
    z2 = oracle.normals(1234, 1, 0, 1 << 28, 500_000)                 # another row: independent stream
    assert abs(np.corrcoef(z, z2)[0, 1]) < 5e-3


def test_exp_accuracy(oracle):
    mp.mp.prec = 120
    rng = np.random.default_rng(3)
    ts = list(-rng.random(1500) * 8) + [0.0, -1e-300, -0.34657359027997264, -0.3465735902799727, -6.68, -50.0, -700.0]
    worst = max(abs(mp.mpf(oracle.exp_neg(float(t))) / mp.exp(mp.mpf(float(t))) - 1) for t in ts)
    assert worst < 4e-16


def _zig_reference(u, s):
    """the definition of the fast path in exact arithmetic, from independently derived tables"""
    from oracle import oracle_np
    W, KH, _, _ = oracle_np.zig_tables()
    i = s & 0x1FF
    x = mp.mpf(u) * mp.mpf(float(W[i])) / 2 ** 32
    return (-x if s >> 9 else x), (u >> 20) < int(KH[i]), i


def test_ziggurat_select_fields(oracle):
    """draw t of a block reads the 10-bit field [sign][layer] from bits 3..12 / 13..22 / {0, 23..31} of the fourth word"""
    from oracle import oracle_np
    rng = np.random.default_rng(11)
    for w in [0, 0xFFFFFFFF, 0x00000008, 0x00001000, 0x00002000, 0x00400000, 0x00800000, 0x80000000, 0x00000001, 0x6] + \
            [int(v) for v in rng.integers(0, 2 ** 32, 200)]:
        bits = [(w >> b) & 1 for b in range(32)]
        want = [sum(bits[3 + b] << b for b in range(10)), sum(bits[13 + b] << b for b in range(10)),
                sum(bits[23 + b] << b for b in range(9)) | (bits[0] << 9)]
        for t in range(3):
            assert oracle.zig_select(w, t) == want[t] == oracle_np.zig_select(w, t)


def test_ziggurat_table_is_a_valid_ziggurat():
    """the generated table: edges strictly decreasing, KH a conservative bound on W'[i+1]/W'[i], every layer's area within
    1e-9 of the common value (the edges are perturbed by < 2^-40 to carry KH in their low bits)"""
    from oracle import oracle_np
    W, KH, F, R = oracle_np.zig_tables()
    assert len(W) == 513 and W[512] == 0.0 and R == W[1] and np.all(np.diff(W) < 0)
    mp.mp.prec = 120
    f = lambda x: mp.exp(-mp.mpf(x) ** 2 / 2)
    V = mp.mpf(R) * f(R) + mp.sqrt(mp.pi / 2) * mp.erfc(mp.mpf(R) / mp.sqrt(2))
    assert abs(mp.mpf(float(W[0])) * f(R) / V - 1) < 1e-9                       # the base strip's virtual width
    for i in range(512):
        assert mp.mpf(int(KH[i])) / 4096 <= mp.mpf(float(W[i + 1])) / mp.mpf(float(W[i]))
        assert int(KH[i]) == (np.float64(W[i]).view(np.uint64) & np.uint64(0xFFF))  # the bound lives in the edge's low bits
        if i >= 1:
            assert abs(mp.mpf(float(W[i])) * (f(W[i + 1]) - f(W[i])) / V - 1) < 1e-9
            assert abs(mp.mpf(float(F[i])) / f(W[i]) - 1) < 2.0 ** -52
    assert abs(R - 3.8520461503683912) < 1e-11          # 512 layers


def test_ziggurat_fast_path_matches_its_definition(oracle):
    mp.mp.prec = 120
    rng = np.random.default_rng(4)
    n_slow = 0
    for _ in range(6000):
        u, s = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 1024))
        z, slow = oracle.zig_from_words(u, s)
        want, ok, _ = _zig_reference(u, s)
        assert slow == (not ok)
        if ok:
            assert abs(mp.mpf(z) - want) <= abs(want) * 2.0 ** -53      # one correctly rounded multiplication
        n_slow += slow
    assert 20 <= n_slow <= 85                                             # 0.81 % of 6000


def test_ziggurat_slow_path_by_layer(oracle):
    """every layer's wedge and the tail, with words chosen to fail the fast test: the result must be finite, carry the
    drawn sign when it is the original candidate, the tail lies beyond R, and the sliver of the base strip between
    the fast bound and R is still accepted as it is"""
    from oracle import oracle_np
    W, KH, _, R = oracle_np.zig_tables()
    for i in range(512):
        for sign in (0, 1):
            s = (sign << 9) | i
            z, slow = oracle.zig_from_words(0xFFFFFFFF, s)            # u = 1 - 2^-32: outermost sliver of the layer
            assert slow and np.isfinite(z)
            if i == 0:
                assert abs(z) > R and (z < 0) == bool(sign)
            want = oracle_np.zig_slow(0xFFFFFFFF, s)
            assert abs(z - want) < 2e-15
    u = (int(KH[0]) << 20) + 5                                            # just past the fast bound of the base strip
    z, slow = oracle.zig_from_words(u, 0)
    assert slow and z == u * (W[0] / 2.0 ** 32) and z < R


def test_extreme_ziggurat_words(oracle):
    for u in (0, 0xffffffff, 0x80000000, 0x7fffffff, 0x000fffff, 0x7f800000):
        for s in (0, 1, 511, 512, 1023):
            z, _ = oracle.zig_from_words(u, s)
            assert np.isfinite(z) and abs(z) < 13.6


def test_ziggurat_normals_are_standard_normal(oracle):
    z = oracle.zig_normals(1234, 0, 0, 1 << 28, 1_000_000)
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs((z ** 4).mean() - 3) < 4 * np.sqrt(96 / n)
    assert stats.kstest(z[:500_000], "norm").pvalue > 1e-3
    # chi-square over 200 equiprobable cells (cuts fall inside layers, wedges and the tail alike)
    cuts = stats.norm.ppf(np.linspace(0, 1, 201)[1:-1])
    cnt = np.bincount(np.searchsorted(cuts, z), minlength=200)
    assert stats.chisquare(cnt).pvalue > 1e-3
    # the tail beyond R and far tails
    R = 3.8520461503683912
    for t in (R, 4.0, 4.5):
        k, p = (np.abs(z) > t).sum(), 2 * stats.norm.sf(t)
        assert abs(k - n * p) < 5 * np.sqrt(n * p) + 1
    for a, b in ((0, 1), (1, 2), (0, 2)):                                   # the three draws of a block
        assert abs(np.corrcoef(z[a::3], z[b::3])[0, 1]) < 4 / np.sqrt(n / 3)
        assert abs(np.corrcoef(np.abs(z[a::3]), np.abs(z[b::3]))[0, 1]) < 4 / np.sqrt(n / 3)
    z2 = oracle.zig_normals(1234, 1, 0, 1 << 28, 1_000_000)                # another row: independent stream
    assert abs(np.corrcoef(z, z2)[0, 1]) < 4 / np.sqrt(n)


def test_acc_uniform_range_and_determinism(oracle):
    u = np.array([oracle.acc_uniform(12, c, it) for c in range(3) for it in range(1, 400)])
    assert (u >= 0).all() and (u < 1).all() and abs(u.mean() - 0.5) < 0.05
    assert oracle.acc_uniform(12, 1, 7) == oracle.acc_uniform(12, 1, 7)
    assert oracle.acc_uniform(12, 1, 7) != oracle.acc_uniform(13, 1, 7)


def test_pair_unrank_is_the_reference_order(oracle):
    # [(i,j) for i in 1:N, j in 1:N if i<j]: i fastest (AlgoBGP.jl:653)
    N = 7
    want = [(i, j) for j in range(N) for i in range(N) if i < j]
    got = [oracle.pair_unrank(q) for q in range(N * (N - 1) // 2)]
    assert got == want
    for q in [0, 1, 2, 523775, 2 ** 25 + 17, 33550335]:
        i, j = oracle.pair_unrank(q)
        assert 0 <= i < j and j * (j - 1) // 2 + i == q


def test_pairs_are_distinct_and_in_range(oracle):
    for N in (2, 3, 5, 64, 300):
        for it in (2, 3, 99):
            p = oracle.pairs(777, it, N)
            assert len(p) == (N - 1 if N < 3 else N)
            assert len({tuple(r) for r in p.tolist()}) == len(p)
            assert (p[:, 0] < p[:, 1]).all() and (p >= 0).all() and (p < N).all()
    assert not np.array_equal(oracle.pairs(777, 2, 64), oracle.pairs(777, 3, 64))
