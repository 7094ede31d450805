"""GPU parity: resamplers and small primitives through the C ABI against the CPU oracle.  Ancestor indices must be
BIT-EXACT given identical weights and uniforms (north_star); floating-point helpers within 1e-12."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.test_oracle_primitives import KINDS, weights  # noqa: E402


@pytest.fixture(scope="module")
def up():
    import unbiased_pimh_b200 as m
    return m


@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 256, 4096, 8192])
def test_systematic_bit_exact(up, oracle, N):
    rng = np.random.default_rng(N)
    B = 24 if N <= 256 else 6
    for kind in KINDS:
        w = np.stack([weights(rng, N, kind) for _ in range(B)])
        u = rng.random(B)
        for n in sorted({1, N, max(1, N // 3)}):
            got = up.systematic_resampling_n_(w, n, u)
            ref = np.stack([oracle.systematic_resampling_n_(w[b], n, float(u[b])) for b in range(B)])
            assert np.array_equal(got, ref), (kind, n)


@pytest.mark.parametrize("N", [1, 2, 31, 33, 256, 4096, 8192])
def test_multinomial_bit_exact(up, oracle, N):
    rng = np.random.default_rng(1000 + N)
    B = 16 if N <= 256 else 4
    for kind in KINDS:
        w = np.stack([weights(rng, N, kind) for _ in range(B)])
        for n in sorted({N, max(1, N // 2)}):
            ur = rng.random((B, n)); us = rng.random((B, max(n - 1, 1)))
            e = np.stack([oracle.sorted_uniforms(ur[b]) for b in range(B)])
            got = up.multinomial_resampling_n_(w, n, e_sorted=e, u_shuffle=us if n > 1 else None)
            ref = np.stack([oracle.multinomial_resampling_n_(w[b], n, ur[b], us[b] if n > 1 else None) for b in range(B)])
            assert np.array_equal(got, ref), (kind, n)
            got0 = up.multinomial_resampling_n_(w, n, e_sorted=e, u_shuffle=None, shuffle=False)
            ref0 = np.stack([oracle.multinomial_from_sorted(w[b], n, e[b], True) for b in range(B)])
            assert np.array_equal(got0, ref0) and np.all(np.diff(got0, axis=1) >= 0)


@pytest.mark.parametrize("N", [1, 2, 64, 1000, 4096])
def test_coupled_multinomial_bit_exact(up, oracle, N):
    rng = np.random.default_rng(2000 + N)
    B = 8
    for kind in ("uniform", "onehot", "zeros", "0.1", "1", "5"):
        w1 = np.stack([weights(rng, N, kind) for _ in range(B)]); w2 = np.stack([weights(rng, N, "1") for _ in range(B)])
        ur = rng.random((B, N)); us = rng.random((B, max(N - 1, 1)))
        e = np.stack([oracle.sorted_uniforms(ur[b]) for b in range(B)])
        got = up.coupled_multinomial_resampling_n_(w1, w2, N, e_sorted=e, u_shuffle=us if N > 1 else None)
        ref = np.stack([oracle.coupled_multinomial_resampling_n_(w1[b], w2[b], N, ur[b], us[b] if N > 1 else None) for b in range(B)])
        assert np.array_equal(got, ref), kind


@pytest.mark.parametrize("N", [2, 50, 257, 4096])
def test_indexmatching_bit_exact(up, oracle, N):
    rng = np.random.default_rng(3000 + N)
    B = 6
    w1 = np.stack([weights(rng, N, "1") for _ in range(B)]); w2 = np.stack([weights(rng, N, "0.1") for _ in range(B)])
    w2[0] = w1[0]                                            # identical systems: alpha == 1 up to rounding
    uc = rng.random((B, N)); um = rng.random((B, N)); ums = rng.random((B, N - 1)); ucm = rng.random((B, N)); ucms = rng.random((B, N - 1))
    e_mult = np.full((B, N), 0.5); e_cm = np.full((B, N), 0.5)
    ref = []
    for b in range(B):
        alpha = 0.0
        for v in np.minimum(w1[b], w2[b]):
            alpha += v
        nc = int((uc[b] < alpha).sum())
        if nc > 0:
            e_mult[b, :nc] = oracle.sorted_uniforms(um[b, :nc])
        if N - nc > 0:
            e_cm[b, :N - nc] = oracle.sorted_uniforms(ucm[b, :N - nc])
        ref.append(oracle.indexmatching(w1[b], w2[b], N, uc[b], um[b], ums[b], ucm[b], ucms[b]))
    got = up.indexmatching_cpp(w1, w2, N, u_couple=uc, e_mult=e_mult, us_mult=ums, e_cm=e_cm, us_cm=ucms)
    assert np.array_equal(got, np.stack(ref))


def test_r_level_wrappers_are_one_based(up):
    rng = np.random.default_rng(1)
    N = 100
    w = weights(rng, N, "1")
    a = up.multinomial_resampling_n(w, N)
    assert a.min() >= 1 and a.max() <= N
    s = up.systematic_resampling_n(w, N, 0.3)
    assert s.min() >= 1 and s.max() <= N and np.all(np.diff(s) >= 0)
    x = np.zeros((1, N))
    for f in (up.CR_systematic, up.CR_multinomial, up.CR_indexmatching):
        c = f(x, x, w, w)
        assert c.shape == (N, 2) and c.min() >= 1 and c.max() <= N
        assert np.array_equal(c[:, 0], c[:, 1])              # identical weights -> fully coupled
    freq = np.bincount(np.concatenate([up.multinomial_resampling_n(w, N) for _ in range(300)]) - 1, minlength=N) / (300 * N)
    assert np.max(np.abs(freq - w) / np.sqrt(w * (1 - w) / (300 * N) + 1e-12)) < 5.5


def test_sorted_uniforms_and_normalize_weight(up, oracle):
    rng = np.random.default_rng(5)
    for n in (1, 2, 33, 4096, 8192):
        u = rng.random((5, n))
        e = up.sorted_uniforms(u)
        ref = np.stack([oracle.sorted_uniforms(u[b]) for b in range(5)])
        assert np.allclose(e, ref, rtol=1e-12, atol=0) and np.all(np.diff(e, axis=1) >= -1e-15)
    for N in (1, 7, 256, 4096):
        lw = rng.normal(size=(9, N)) * 20 - 300
        w, lz = up.normalize_weight(lw, return_logz=True)
        for b in range(9):
            lzr, wr = oracle.logz_increment(lw[b])
            assert np.allclose(w[b], wr, rtol=1e-13, atol=0) and abs(lz[b] - lzr) <= 1e-12 * abs(lzr)   # fp64, tolerance 1e-12 relative


def test_mvnorm_family(up, oracle):
    rng = np.random.default_rng(6)
    for d in (1, 3, 10, 64):
        A = rng.normal(size=(d, d)); cov = A @ A.T + d * np.eye(d); L = np.linalg.cholesky(cov); mean = rng.normal(size=d)
        x = rng.normal(size=(d, 500)) * 2
        for got, ref in ((up.fast_dmvnorm_transpose_cholesky(x, mean, L), oracle.dmvnorm_transpose_cholesky(x, mean, L)),
                         (up.fast_dmvnorm_transpose(x, mean, cov), oracle.dmvnorm_transpose(x, mean, cov)),
                         (up.fast_dmvnorm(x.T.copy(), mean, cov), oracle.dmvnorm(x.T.copy(), mean, cov))):
            assert np.allclose(got, ref, rtol=1e-11, atol=1e-9)                                         # fp64, 1e-11 relative
        z = rng.normal(size=(d, 200))
        assert np.allclose(up.fast_rmvnorm_transpose_cholesky(200, mean, L, z=z), oracle.rmvnorm_transpose_cholesky(200, mean, L, z), rtol=1e-12, atol=1e-12)
        assert np.allclose(up.fast_rmvnorm_transpose(200, mean, cov, z=z), oracle.rmvnorm_transpose(200, mean, cov, z), rtol=1e-11, atol=1e-11)
        zn = rng.normal(size=(200, d))
        assert np.allclose(up.fast_rmvnorm(200, mean, cov, z=zn), oracle.rmvnorm(200, mean, cov, zn), rtol=1e-11, atol=1e-11)
    y = up.fast_rmvnorm_transpose(40000, np.array([1.0, -2.0]), np.array([[2.0, 0.6], [0.6, 1.0]]), seed=3)   # Philox normals
    assert np.allclose(y.mean(axis=1), [1, -2], atol=0.03) and np.allclose(np.cov(y), [[2, 0.6], [0.6, 1]], atol=0.05)
    assert np.array_equal(up.create_A(0.5, 6), oracle.create_A(0.5, 6))


def test_tree_object_matches_oracle_tree(up, oracle):
    """class Tree (src/tree.cpp): same slot numbering, offspring counts, leaves and paths as the sequential algorithm,
    including growth through double_size()."""
    rng = np.random.default_rng(7)
    for (N, d, M) in ((1, 1, 0), (8, 2, 0), (100, 3, 0), (1000, 1, 0), (300, 2, 6000)):
        tg, to = up.Tree(N, M, d), oracle.Tree(N, M, d)
        x = rng.normal(size=(d, N)); tg.init(x); to.init(x)
        for t in range(40):
            w = weights(rng, N, "0.1" if t % 2 else "1")
            a = np.sort(rng.choice(N, size=N, p=w)).astype(np.int32)
            if t % 4 == 3:
                a = rng.permutation(a).astype(np.int32)
            x = rng.normal(size=(d, N))
            tg.update(x, a); to.update(x, a)
        assert (tg.N, tg.M, tg.dimx, tg.nsteps) == (N, to.M, d, 40)
        assert np.array_equal(tg.l_star, to.l_star) and np.array_equal(tg.o_star, to.o_star)
        live = to.o_star > 0
        live[to.l_star] = True
        assert np.array_equal(tg.a_star[live], to.a_star[live])
        for n in range(0, N, max(1, N // 7)):
            assert np.array_equal(tg.get_path(n), to.get_path(n))
        for lag in (0, 3):
            assert np.array_equal(tg.retrieve_xgeneration(lag), to.retrieve_xgeneration(lag))
        tg.close()


def test_rng_specification_functions_are_bit_identical_on_the_gpu(oracle):
    """rng_log / rng_sincos2pi: the same double on the CPU oracle and on the B200 for every uniform (edge cases included)."""
    from unbiased_pimh_b200 import rng as prng
    g = np.random.default_rng(6)
    u = np.concatenate([g.random(300000), 1.0 - g.random(50000) * 1e-9, g.random(50000) * 1e-12, (np.arange(1, 2049) - 0.5) / 2048,
                        [2.0 ** -53, 1.0 - 2.0 ** -53, 0.5, 0.25, 0.75, 0.125, 0.375, 0.7071067811865476, 0.7071067811865475]])
    u = u[(u > 0) & (u < 1)]
    ol, os_, oc = oracle.rng_transform(u)
    gl, gs, gc = prng.rng_transform(u)
    assert np.array_equal(ol.view(np.int64), gl.view(np.int64))
    assert np.array_equal(os_.view(np.int64), gs.view(np.int64)) and np.array_equal(oc.view(np.int64), gc.view(np.int64))
