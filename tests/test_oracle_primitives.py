"""CPU tests of the oracle's primitives against independent brute-force restatements written here in numpy/Python
(the reference ships no tests or golden vectors for them -- SURVEY 4 -- so they are pinned by re-derivation)."""
import math

import numpy as np
import pytest


# ---------------------------------------------------------------- Philox
def test_philox_known_answers(oracle):
    # Random123 kat_vectors: philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_uniforms_open_interval(oracle):
    u = np.array([oracle.philox_uniform2(17, i, 3, 5, 1) for i in range(2000)]).ravel()
    assert u.min() > 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1 / 12) < 0.01


# ---------------------------------------------------------------- resamplers
def weights(rng, N, kind):
    if kind == "uniform":
        w = np.full(N, 1.0 / N)
    elif kind == "onehot":
        w = np.zeros(N); w[rng.integers(N)] = 1.0
    elif kind == "zeros":
        w = rng.random(N); w[rng.random(N) < 0.5] = 0.0
        if w.sum() == 0: w[0] = 1.0
        w /= w.sum()
    elif kind == "ties":
        w = np.round(rng.random(N) * 4 + 1); w /= w.sum()
    elif kind == "tiny":
        w = np.exp(rng.normal(size=N) * 25); w /= w.sum()
    else:
        s = float(kind)
        w = np.exp(rng.normal(size=N) * s); w /= w.sum()
    return w


KINDS = ["uniform", "onehot", "zeros", "ties", "tiny", "0.1", "1", "5"]


def ref_systematic(w, n, u):
    """src/systematic.cpp:7-32 transcribed statement by statement (Python floats are IEEE doubles)."""
    anc = np.empty(n, dtype=np.int32)
    u = u / n
    j = 0
    csw = w[0]
    for k in range(n):
        while csw < u and j < len(w) - 1:
            j += 1
            csw += w[j]
        u = u + 1.0 / n
        anc[k] = j
    return anc


def ref_multinomial_sorted(w, n, e, scale):
    cdf = np.empty(len(w)); acc = 0.0
    for i, x in enumerate(w):
        acc = x if i == 0 else acc + x
        cdf[i] = acc
    sumw = cdf[-1]
    anc = np.empty(n, dtype=np.int32); j = len(w)
    for i in range(n, 0, -1):
        u = sumw * e[i - 1] if scale else e[i - 1]
        while j > 0 and u < cdf[j - 1]:
            j -= 1
        anc[i - 1] = min(j, len(w) - 1)
    return anc


def ref_shuffle(a, us):
    a = list(a)
    for i in range(1, len(a)):
        j = min(int(math.floor(us[i - 1] * (i + 1))), i)
        a[i], a[j] = a[j], a[i]
    return np.array(a, dtype=np.int32)


@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 256, 1000])
def test_systematic_matches_transcription(oracle, N):
    rng = np.random.default_rng(N)
    for kind in KINDS:
        w = weights(rng, N, kind)
        for n in (1, N, max(1, N // 3)):
            u = float(rng.random())
            assert np.array_equal(oracle.systematic_resampling_n_(w, n, u), ref_systematic(w, n, u)), (kind, n)


@pytest.mark.parametrize("N", [1, 2, 33, 256, 777])
def test_multinomial_and_shuffle_match_transcription(oracle, N):
    rng = np.random.default_rng(100 + N)
    for kind in KINDS:
        w = weights(rng, N, kind)
        for n in (N, max(1, N // 2), 2 * N):
            ur = rng.random(n); us = rng.random(max(n - 1, 1))
            e = oracle.sorted_uniforms(ur)
            # recurrence of src/multinomial.cpp:20-24
            ln = 0.0; e_ref = np.empty(n)
            for i in range(n, 0, -1):
                ln += math.log(ur[i - 1]) / i
                e_ref[i - 1] = math.exp(ln)
            assert np.array_equal(e, e_ref)
            assert np.all(np.diff(e) >= 0)
            pre = ref_multinomial_sorted(w, n, e, True)
            assert np.array_equal(oracle.multinomial_from_sorted(w, n, e, True), pre)
            assert np.array_equal(oracle.multinomial_resampling_n_(w, n, ur, us if n > 1 else None), ref_shuffle(pre, us) if n > 1 else pre)


def test_multinomial_is_unbiased(oracle):
    rng = np.random.default_rng(3)
    N = 16
    w = weights(rng, N, "1")
    counts = np.zeros(N)
    reps = 4000
    for _ in range(reps):
        a = oracle.multinomial_resampling_n_(w, N, rng.random(N), rng.random(N - 1))
        counts += np.bincount(a, minlength=N)
    p = counts / (reps * N)
    assert np.max(np.abs(p - w) / np.sqrt(w * (1 - w) / (reps * N) + 1e-12)) < 5.0


@pytest.mark.parametrize("N", [1, 2, 64, 301])
def test_coupled_multinomial_matches_transcription(oracle, N):
    rng = np.random.default_rng(7 + N)
    for kind in ("0.1", "1", "zeros", "onehot"):
        w1, w2 = weights(rng, N, kind), weights(rng, N, "1")
        n = N
        ur = rng.random(n); us = rng.random(max(n - 1, 1))
        e = oracle.sorted_uniforms(ur)
        a1 = ref_multinomial_sorted(w1, n, e, False); a2 = ref_multinomial_sorted(w2, n, e, False)
        idx = ref_shuffle(np.arange(n), us) if n > 1 else np.arange(n)
        got = oracle.coupled_multinomial_resampling_n_(w1, w2, n, ur, us if n > 1 else None)
        assert np.array_equal(got[:, 0], a1[idx]) and np.array_equal(got[:, 1], a2[idx])
        # identical weights => identical columns
        same = oracle.coupled_multinomial_resampling_n_(w1, w1, n, ur, us if n > 1 else None)
        assert np.array_equal(same[:, 0], same[:, 1])


@pytest.mark.parametrize("N", [2, 50, 257])
def test_indexmatching_matches_transcription(oracle, N):
    rng = np.random.default_rng(11 + N)
    for trial in range(6):
        w1, w2 = weights(rng, N, "1"), weights(rng, N, "1")
        if trial == 0:
            w2 = w1.copy()                          # alpha == 1 up to rounding: everything coupled
        n = N
        uc = rng.random(N); um, ums, ucm, ucms = rng.random(n), rng.random(n), rng.random(n), rng.random(n)
        got = oracle.indexmatching(w1, w2, n, uc, um, ums, ucm, ucms)
        nu = np.minimum(w1, w2); alpha = 0.0
        for v in nu: alpha += v
        mu = nu / alpha
        with np.errstate(divide="ignore", invalid="ignore"):
            R1, R2 = (w1 - nu) / (1 - alpha), (w2 - nu) / (1 - alpha)
        cpl = uc[:n] < alpha; nc = int(cpl.sum())
        a = oracle.multinomial_resampling_n_(mu, nc, um[:nc], ums[:max(nc - 1, 1)] if nc > 1 else None) if nc > 0 else None
        aR = oracle.coupled_multinomial_resampling_n_(R1, R2, n - nc, ucm[:n - nc], ucms[:max(n - nc - 1, 1)] if n - nc > 1 else None) if nc < n else None
        exp = np.empty((n, 2), dtype=np.int32); ci = ui = 0
        for i in range(n):
            if cpl[i]: exp[i] = a[ci]; ci += 1
            else: exp[i] = aR[ui]; ui += 1
        assert np.array_equal(got, exp)
        assert np.all(got[cpl, 0] == got[cpl, 1])


def test_normalize_weight_and_logz(oracle):
    rng = np.random.default_rng(5)
    for N in (1, 7, 1000):
        lw = rng.normal(size=N) * 30 - 500
        w = oracle.normalize_weight(lw)
        m = lw.max(); ref = np.exp(lw - m); ref = ref / float(np.sum(ref.astype(np.longdouble)))
        assert np.allclose(w, ref, rtol=1e-15, atol=0) and abs(w.sum() - 1) < 1e-12
        lz, w2 = oracle.logz_increment(lw)
        assert np.array_equal(w, w2)
        assert abs(lz - (math.log(np.mean(np.exp(lw - m).astype(np.longdouble))) + m)) < 1e-12 * abs(lz)


# ---------------------------------------------------------------- tree
@pytest.mark.parametrize("N,d,M", [(1, 1, 0), (8, 2, 0), (64, 3, 0), (200, 1, 0), (50, 2, 5000)])
def test_tree_paths_equal_dense_ancestry(oracle, N, d, M):
    """Tree.get_path must return exactly the path read off a dense (T+1) x N store of states and ancestors;
    M=0 starts at the minimum capacity 3N so double_size (src/tree.cpp:101-116) is exercised."""
    rng = np.random.default_rng(N * 31 + d)
    T = 60
    tree = oracle.Tree(N, M, d)
    x = rng.normal(size=(d, N)); tree.init(x)
    hist_x, hist_a = [x.copy()], []
    for t in range(T):
        w = weights(rng, N, "0.1" if t % 3 else "1")
        a = np.sort(rng.choice(N, size=N, p=w)).astype(np.int32)
        if t % 5 == 4:
            a = rng.permutation(a).astype(np.int32)
        x = rng.normal(size=(d, N))
        tree.update(x, a)
        hist_x.append(x.copy()); hist_a.append(a)
    assert tree.nsteps == T
    if M == 0 and N >= 8:
        assert tree.M > 3 * N                 # it grew
    for n in range(N):
        path = tree.get_path(n)
        j = n
        for step in range(T, -1, -1):
            assert np.array_equal(path[:, step], hist_x[step][:, j])
            if step > 0:
                j = hist_a[step - 1][j]
    for lag in (0, 1, 7):
        xg = tree.retrieve_xgeneration(lag)
        for n in range(0, N, max(1, N // 5)):
            j = n
            for s in range(lag):
                j = hist_a[T - 1 - s][j]
            assert np.array_equal(xg[:, n], hist_x[T - lag][:, j])
    # bookkeeping invariants of prune/insert (src/tree.cpp:45-53, 70-86)
    a_star, o_star, l_star = tree.a_star, tree.o_star, tree.l_star
    assert len(set(l_star.tolist())) == N and np.all(o_star >= 0)
    assert np.all(o_star[l_star] == 0)


# ---------------------------------------------------------------- Gaussian kernels
def test_dmvnorm_family_against_scipy(oracle):
    from scipy.stats import multivariate_normal
    rng = np.random.default_rng(2)
    for d in (1, 3, 10):
        A = rng.normal(size=(d, d)); cov = A @ A.T + d * np.eye(d)
        mean = rng.normal(size=d); x = rng.normal(size=(d, 40)) * 2
        ref = multivariate_normal(mean, cov).logpdf(x.T)
        shift = d * (0.5 * math.log(2 * math.pi) - 0.9189385)          # the reference's truncated constant, src/mvnorm.cpp:121
        L = np.linalg.cholesky(cov)
        for got in (oracle.dmvnorm_transpose_cholesky(x, mean, L), oracle.dmvnorm_transpose(x, mean, cov), oracle.dmvnorm(x.T.copy(), mean, cov)):
            assert np.allclose(got - shift, ref, rtol=1e-12, atol=1e-10)


def test_rmvnorm_family(oracle):
    rng = np.random.default_rng(4)
    d, n = 4, 9
    A = rng.normal(size=(d, d)); cov = A @ A.T + d * np.eye(d); L = np.linalg.cholesky(cov); mean = rng.normal(size=d)
    z = rng.normal(size=(d, n))
    assert np.allclose(oracle.rmvnorm_transpose_cholesky(n, mean, L, z), L @ z + mean[:, None], rtol=1e-13)
    assert np.allclose(oracle.rmvnorm_transpose(n, mean, cov, z), L @ z + mean[:, None], rtol=1e-12)
    zn = rng.normal(size=(n, d))
    assert np.allclose(oracle.rmvnorm(n, mean, cov, zn), zn @ L.T + mean[None, :], rtol=1e-12)
    assert np.allclose(oracle.create_A(0.5, 4), [[0.5 ** (1 + abs(i - j)) for j in range(4)] for i in range(4)])


def test_rng_specification_log_and_sincos(oracle):
    """oracle/rngmath.h: the fixed-operation-sequence log and sin/cos(2 pi u) of the RNG specification are 1-ulp accurate."""
    rng = np.random.default_rng(5)
    u = np.concatenate([rng.random(200000), 1.0 - rng.random(50000) * 1e-9, rng.random(50000) * 1e-12,
                        [2.0 ** -53, 1.0 - 2.0 ** -53, 0.5, 0.25, 0.75, 0.125, 0.7071067811865476, 0.7071067811865475]])
    u = u[(u > 0) & (u < 1)]
    lg, sn, cs = oracle.rng_transform(u)
    assert np.max(np.abs(lg - np.log(u)) / np.abs(np.log(u))) < 4.5e-16
    ang = 2 * np.pi * u.astype(np.longdouble)
    assert np.max(np.abs(sn - np.sin(ang).astype(np.float64))) < 4e-16 and np.max(np.abs(cs - np.cos(ang).astype(np.float64))) < 4e-16
    assert np.max(np.abs(sn * sn + cs * cs - 1)) < 1e-15
    # Box-Muller on top of them gives standard normals (moments and a KS test)
    from scipy import stats
    ua, ub = rng.random(400000), rng.random(400000)
    la, _, _ = oracle.rng_transform(ua)
    _, s2, c2 = oracle.rng_transform(ub)
    z = np.concatenate([np.sqrt(-2 * la) * c2, np.sqrt(-2 * la) * s2])
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1) < 6e-3 and abs(stats.kurtosis(z)) < 0.02
    assert stats.kstest(z[:100000], "norm").pvalue > 1e-3
