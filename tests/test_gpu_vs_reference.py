"""The CUDA natives (through the C ABI) against the REFERENCE'S OWN CODE: oracle/_ref/libref.so is /root/reference/src/
{systematic,multinomial,coupledresamplingindexmatching,tree}.cpp compiled unmodified (oracle/ref_build/); it travels to the GPU
box with the snapshot.  Ancestors, slot arrays and paths must be identical, bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests.test_oracle_primitives import KINDS, weights  # noqa: E402

refbind = pytest.importorskip("oracle.refbind")
if not os.path.exists(refbind.LIB_PATH):
    pytest.skip("oracle/_ref/libref.so not present (built where /root/reference is available)", allow_module_level=True)


@pytest.fixture(scope="module")
def up():
    import unbiased_pimh_b200 as m
    return m


@pytest.mark.parametrize("N", [1, 2, 33, 256, 4096])
def test_systematic_equals_reference_binary(up, N):
    rng = np.random.default_rng(N)
    for kind in KINDS:
        w = weights(rng, N, kind)
        for n in sorted({1, N, max(1, N // 3)}):
            u = float(rng.random())
            a_ref, oob = refbind.systematic_resampling_n_(w, n, u)
            if oob:
                continue
            assert np.array_equal(up.systematic_resampling_n_(w, n, u), a_ref), (kind, n)


@pytest.mark.parametrize("N", [2, 33, 256, 4096])
def test_multinomial_and_coupled_equal_reference_binary(up, oracle, N):
    rng = np.random.default_rng(10 + N)
    for kind in KINDS:
        w = weights(rng, N, kind); w2 = weights(rng, N, "1")
        for n in sorted({N, max(2, N // 2)}):
            ur = rng.random(n); us = rng.random(n - 1)
            a_ref, oob = refbind.multinomial_resampling_n_(w, n, ur, us)
            assert oob == 0
            e = oracle.sorted_uniforms(ur)        # the recurrence of src/multinomial.cpp:22-24 (pinned to the reference on CPU)
            assert np.array_equal(up.multinomial_resampling_n_(w, n, e_sorted=e, u_shuffle=us), a_ref), (kind, n)
            c_ref, _ = refbind.coupled_multinomial_resampling_n_(w, w2, n, ur, us)
            got = up.coupled_multinomial_resampling_n_(w, w2, n, e_sorted=e, u_shuffle=us)
            assert np.array_equal(got, np.minimum(c_ref, N - 1)), (kind, n)      # the reference can return N when sum(w) < e (H8): clamped here


@pytest.mark.parametrize("N", [2, 64, 1000])
def test_indexmatching_equals_reference_binary(up, oracle, N):
    rng = np.random.default_rng(20 + N)
    for k1 in ("0.1", "1", "5", "zeros"):
        w1, w2 = weights(rng, N, k1), weights(rng, N, "1")
        n = N
        uc = rng.random(N)
        alpha = 0.0
        for a, b in zip(w1, w2):
            alpha += min(a, b)
        nc = int(np.sum(uc[:n] < alpha)); nres = n - nc
        um, ums = rng.random(nc), rng.random(max(nc - 1, 0))
        ur, urs = rng.random(nres), rng.random(max(nres - 1, 0))
        a_ref, oob = refbind.indexmatching(w1, w2, n, np.concatenate([uc, um, ums, ur, urs]))
        e_m = np.full(n, 0.5); e_c = np.full(n, 0.5)
        if nc:
            e_m[:nc] = oracle.sorted_uniforms(um)
        if nres:
            e_c[:nres] = oracle.sorted_uniforms(ur)
        us_m = np.full(n - 1, 0.5); us_c = np.full(n - 1, 0.5)
        us_m[:max(nc - 1, 0)] = ums; us_c[:max(nres - 1, 0)] = urs
        got = up.indexmatching_cpp(w1, w2, n, u_couple=uc, e_mult=e_m, us_mult=us_m, e_cm=e_c, us_cm=us_c)
        assert np.array_equal(got, np.minimum(a_ref, N - 1)), k1


@pytest.mark.parametrize("N,d", [(64, 2), (256, 1), (1000, 4)])
def test_tree_equals_reference_binary(up, N, d):
    rng = np.random.default_rng(N + d)
    tg, tr = up.Tree(N, 0, d), refbind.Tree(N, 0, d)
    x0 = rng.normal(size=(d, N)); tg.init(x0); tr.init(x0)
    for step in range(30):
        if step % 10 == 7:
            a = np.arange(N, dtype=np.int32)
        elif step % 10 == 3:
            a = np.full(N, rng.integers(N), dtype=np.int32)
        else:
            w = np.exp(rng.normal(size=N) * (0.3 if step % 2 else 2.0)); w /= w.sum()
            a = np.sort(rng.choice(N, size=N, p=w)).astype(np.int32)
        x = rng.normal(size=(d, N))
        tg.update(x, a); tr.update(x, a)
        ar, orr, lr, xr = tr.arrays()
        assert tg.M == tr.M and np.array_equal(tg.l_star, lr) and np.array_equal(tg.a_star, ar) and np.array_equal(tg.o_star, orr), step
    for n in (0, N // 2, N - 1):
        assert np.array_equal(tg.get_path(n), tr.get_path(n))
    assert np.array_equal(tg.retrieve_xgeneration(3), tr.retrieve_xgeneration(3))
    tg.close()
