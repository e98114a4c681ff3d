"""The float filter of the count kernels (coexpression_b200/csrc/count_fine.cuh: cf_delta, count_dedup.cuh: dd_delta), restated in
numpy float32 arithmetic: a streamed null value and a statistic are the SAME five-rounding float expression of an exact integer sum
and two norms, so each is within 5 * 2^-24 of the reference's double, and a float difference of at least 7e-7 |t| decides the order
of the doubles.  (The GPU tests compare every resulting bisect index with the oracle; this pins the error budget itself.)"""
import numpy as np


def _float_value(S, sxu, sxc):
    """fl(fl(int -> float) * fl(fl(0.5 / sx_u) * fl(0.5 / sx_c))), the kernels' expression (no fused operation in it)."""
    iu = np.float32(0.5 / sxu)
    w = (0.5 / sxc).astype(np.float32)
    return S.astype(np.float32) * (iu * w)


def _double_value(S, sxu, sxc):
    """the reference on rank rows: xy / (sqrt(xx) * sqrt(yy)) with xy = S / 4 (coexpression_v2.c:113)"""
    return (0.25 * S.astype(np.float64)) / (sxu * sxc)


def _rows(rng, n, count):
    # half-norms of rank rows: sqrt(sum u^2) / 2 with u = 2 rank - (n + 1); ties lower the sum
    suu_max = n * (n * n - 1) / 3.0
    suu = np.floor(rng.uniform(0.05, 1.0, size=count) * suu_max)
    sx = np.sqrt(suu) / 2.0
    return sx


def test_float_value_is_within_five_roundings_of_the_double():
    rng = np.random.default_rng(1)
    for n in (24, 256, 1829):
        sxc = _rows(rng, n, 200000)
        sxu = float(_rows(rng, n, 1)[0])
        bound = 4.0 * sxu * sxc                                   # |S| <= 4 sx_u sx_c (Cauchy-Schwarz)
        S = np.rint(rng.uniform(-1, 1, size=sxc.size) * bound).astype(np.int64)
        S = np.clip(S, -(2 ** 31) + 1, 2 ** 31 - 1)
        v32 = _float_value(S, sxu, sxc).astype(np.float64)
        v64 = _double_value(S, sxu, sxc)
        nz = v64 != 0
        rel = np.abs(v32[nz] - v64[nz]) / np.abs(v64[nz])
        assert rel.max() <= 5 * 2.0 ** -24 * (1 + 1e-6), (n, rel.max())
        assert np.all(v32[~nz] == 0)                               # an exact zero stays an exact zero


def test_a_certain_float_compare_never_contradicts_the_doubles():
    rng = np.random.default_rng(2)
    delta = np.float32(7e-7)
    for n in (256, 1829):
        sxc = _rows(rng, n, 400000)
        sxu = float(_rows(rng, n, 1)[0])
        # values crowded around a few statistics: integer sums a handful of units apart
        centre = np.rint(rng.uniform(-0.3, 0.3, size=sxc.size) * 4.0 * sxu * sxc)
        S = (centre + rng.integers(-3, 4, size=sxc.size)).astype(np.int64)
        v32 = _float_value(S, sxu, sxc)
        v64 = _double_value(S, sxu, sxc)
        order = np.argsort(v64, kind="stable")
        a, b = order[:-1], order[1:]                               # neighbours in the exact order: the hardest pairs
        for lo, hi in ((a, b), (rng.permutation(a), rng.permutation(b))):
            t32, t64 = v32[hi], v64[hi]
            d = v32[lo] - t32
            certain = np.abs(d) >= delta * np.abs(t32)
            below32 = d < 0
            below64 = v64[lo] < t64
            assert np.array_equal(below32[certain], below64[certain])
        # and the filter is not vacuous: nearly every pair IS certain
        d = v32[a] - v32[b]
        assert np.mean(np.abs(d) >= delta * np.abs(v32[b])) > 0.5


def test_digit_scheme_of_the_tensor_core_contraction_is_exact():
    """The Spearman contraction in integers (coexpression_b200/csrc/prepare.cuh: rank_pack_kernel, gemm_tc.cuh): u = 2 rank - (n + 1)
    = 128 hi + lo with lo in [-64, 63] and hi an int8 for every row length the library accepts (n <= 8192), the four int8 digit
    products recombine to the exact dot product, S fits int32 up to n = 1860 (the guard in coex_network_batch), and the digit-split
    CTA pair may form its two brackets modulo 2^32 (gemm_i8_tc_split_kernel)."""
    from scipy.stats import rankdata
    rng = np.random.default_rng(3)
    for n in (2, 37, 256, 1829, 1860, 8192):
        X = rng.gamma(0.5, 1.0, size=(6, n))
        X[X < 0.4] = 0.0                                          # ties
        X[0] = np.arange(n)                                       # an untied row: the largest possible |u| and sum of squares
        X[1] = X[0]                                               # ... against itself: the largest possible S
        X[2] = -X[0]
        u = np.rint(2.0 * rankdata(X, axis=1) - (n + 1)).astype(np.int64)
        hi = (u + 64) >> 7                                        # floor((u + 64) / 128): arithmetic shift, as in the kernel
        lo = u - (hi << 7)
        assert lo.min() >= -64 and lo.max() <= 63
        assert hi.min() >= -128 and hi.max() <= 127
        assert np.array_equal(128 * hi + lo, u)
        HH, HL, LH, LL = hi @ hi.T, hi @ lo.T, lo @ hi.T, lo @ lo.T
        S = u @ u.T
        assert np.array_equal(16384 * HH + 128 * (HL + LH) + LL, S)
        if n <= 1860:
            assert np.abs(S).max() < 2 ** 31                      # (n = 1860: 2 144 865 540 < 2 147 483 648)
            assert np.abs(np.stack([HH, HL, LH, LL])).max() < 2 ** 31          # the int32 accumulators of the MMAs
            leader = (16384 * HH + 128 * HL) & 0xFFFFFFFF         # each bracket may wrap ...
            peer = (128 * LH + LL) & 0xFFFFFFFF
            both = (leader + peer) & 0xFFFFFFFF
            assert np.array_equal(np.where(both >= 2 ** 31, both - 2 ** 32, both), S)   # ... their sum is S
        else:
            assert np.abs(S).max() >= 2 ** 31                     # why longer rows are refused for the int32 strip


def _bucket_model(x32, nb, r1):
    """A monotone piecewise-linear bucket map like cf_bucket (three quarters of the buckets for |x| <= r1); the scheme below is
    exact for ANY non-decreasing map, which is all the kernel's float version has to be."""
    x = x32.astype(np.float64)
    st = 0.125 * nb / (1.0 - r1)
    k = 0.375 * nb / r1 - st
    b = np.floor(nb / 2 + st * x + k * np.clip(x, -r1, r1))
    return np.clip(b, 0, nb - 1).astype(np.int64)


def test_bucket_table_counting_scheme_is_exact():
    """count_fine_kernel restated on the host: #values below a statistic = prefix of the bucket table + the certain float compares
    and the exact double compares of its own bucket, corrected by the values of the neighbouring buckets that are closer than the
    float uncertainty -- equal to counting with the reference's doubles, for tiny tables (everything shares buckets), heavy ties
    (equal integer sums and norms) and statistics next to bucket edges."""
    rng = np.random.default_rng(4)
    delta = np.float32(7e-7)
    for trial, (N, n, nb) in enumerate([(3000, 256, 64), (3000, 256, 1024), (5000, 1829, 4096), (4000, 37, 16)]):
        sx = _rows(rng, n, N)
        sx[rng.random(N) < 0.3] = sx[0]                            # many columns share a norm ...
        sxu = float(sx[1])
        S = np.rint(rng.normal(0, 0.08, size=N) * 4.0 * sxu * sx).astype(np.int64)
        S[rng.random(N) < 0.3] = S[2]                              # ... and an integer sum: exact ties
        S[rng.random(N) < 0.02] = 0
        active = rng.random(N) > 0.01                              # zero-variance columns are counted in bulk elsewhere
        v32 = _float_value(S, sxu, sx)
        r64 = _double_value(S, sxu, sx)
        r1 = min(0.5, 6.0 / np.sqrt(n))
        bv = _bucket_model(v32, nb, r1)
        table = np.bincount(bv[active], minlength=nb)
        prefix = np.concatenate([[0], np.cumsum(table)[:-1]])
        cols = rng.choice(np.flatnonzero(active), size=300, replace=False)
        for j in cols:
            t32, tj = v32[j], r64[j]
            dl = delta * np.abs(t32)
            hb = _bucket_model(np.float32([t32]), nb, r1)[0]
            b0 = _bucket_model(np.float32([t32 - dl]), nb, r1)[0]
            b1 = _bucket_model(np.float32([t32 + dl]), nb, r1)[0]
            cnt = int(prefix[hb])
            for b in range(b0, b1 + 1):
                members = np.flatnonzero(active & (bv == b))
                d = v32[members] - t32
                certain = np.abs(d) >= dl
                less_exact = r64[members] < tj
                if b == hb:
                    cnt += int(np.sum(certain & (d < 0))) + int(np.sum(~certain & less_exact))
                elif b < hb:                                       # the prefix has counted them as below
                    cnt -= int(np.sum(~certain & ~less_exact))
                else:                                              # the prefix has not
                    cnt += int(np.sum(~certain & less_exact))
            assert cnt == int(np.sum(active & (r64 < tj))), (trial, int(j))


def test_sorted_statistics_kernel_filter_rule():
    """count_dedup.cuh (dd_delta): there the statistic is ONE rounding of its double (thr32 = (float) t) and the window is
    relative to the streamed value: |v32 - v64| <= 5 * 2^-24 |v|, |t32 - t64| <= 2^-24 |t|; a float difference of at least
    5e-7 |v32| (floor 1e-30) decides the order of the doubles, for both neighbours of v among the statistics."""
    rng = np.random.default_rng(7)
    for n in (256, 1829):
        sxc = _rows(rng, n, 400000)
        sxu = float(_rows(rng, n, 1)[0])
        centre = np.rint(rng.uniform(-0.4, 0.4, size=sxc.size) * 4.0 * sxu * sxc)
        S = (centre + rng.integers(-3, 4, size=sxc.size)).astype(np.int64)
        v32 = _float_value(S, sxu, sxc)
        v64 = _double_value(S, sxu, sxc)
        t32_all = v64.astype(np.float32)                           # a statistic IS some column's double, rounded once
        order = np.argsort(v64, kind="stable")
        a, b = order[:-1], order[1:]
        n_certain = 0
        for val, stat in ((a, b), (b, a), (rng.permutation(a), rng.permutation(b))):
            d = v32[val] - t32_all[stat]
            window = np.maximum(np.float32(5e-7) * np.abs(v32[val]), np.float32(1e-30))
            certain = np.abs(d) >= window
            assert np.array_equal((d < 0)[certain], (v64[val] < v64[stat])[certain])
            assert np.array_equal((d > 0)[certain], (v64[val] > v64[stat])[certain])
            n_certain += int(certain.sum())
        assert n_certain > 0.5 * 3 * a.size
        # exact ties (same integer sum, same norm) are never "certain": they go to the double refinement
        same = v32 - v64.astype(np.float32)
        assert np.all(np.abs(same) < np.maximum(np.float32(5e-7) * np.abs(v32), np.float32(1e-30)) + (v32 == 0))
