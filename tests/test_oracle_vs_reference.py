"""Pins the CPU oracle (oracle/*.c) to the REFERENCE'S OWN CODE: oracle/_ref/libref.so is /root/reference/src/{systematic,
multinomial,coupledresamplingindexmatching,tree,create_A_,SVLevy_functions}.cpp compiled byte for byte against a stand-in
Rcpp.h (oracle/ref_build/).  R's random numbers are the only thing replaced: the tests queue uniforms in the order the
reference consumes them and feed the oracle the same numbers.

Everything here is bit-exact (integer ancestors, slot arrays; doubles compared with ==).  CPU only."""
import numpy as np
import pytest

from tests.test_oracle_primitives import KINDS, weights

refbind = pytest.importorskip("oracle.refbind")
pytestmark = pytest.mark.skipif(not refbind.available(), reason="neither oracle/_ref/libref.so nor /root/reference is present")


@pytest.fixture(scope="module")
def ref():
    refbind.build()
    return refbind


def test_reference_sources_are_the_unmodified_files(ref):
    """The recipe records the sha256 of the files it compiled; when the reference tree is present they must still match."""
    import hashlib, os
    sums = os.path.join(os.path.dirname(ref.LIB_PATH), "SOURCES.sha256")
    assert os.path.exists(sums)
    if not os.path.isdir(os.path.join(ref.REFERENCE, "src")):
        pytest.skip("reference tree not present (GPU box): prebuilt libref.so in use")
    for line in open(sums):
        digest, name = line.split()
        assert hashlib.sha256(open(os.path.join(ref.REFERENCE, "src", name), "rb").read()).hexdigest() == digest, name


# ---------------------------------------------------------------- a4: src/systematic.cpp:7-32
@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 256, 1000, 4096])
def test_systematic_equals_reference(oracle, ref, N):
    rng = np.random.default_rng(N)
    checked = 0
    for kind in KINDS:
        w = weights(rng, N, kind)
        for n in (1, N, max(1, N // 3), 2 * N + 1):
            for u in (float(rng.random()), 1e-300, 0.999999):
                a_ref, oob = ref.systematic_resampling_n_(w, n, u)
                a_orc = oracle.systematic_resampling_n_(w, n, u)
                if oob:       # sum(w) < u: the reference reads past the end; the oracle clamps at N-1 (H8)
                    ok = a_ref < N
                    assert np.array_equal(a_orc[ok], a_ref[ok]) and np.all(a_orc[~ok] == N - 1)
                else:
                    assert np.array_equal(a_orc, a_ref), (kind, n, u)
                    checked += 1
    assert checked > 40


# ---------------------------------------------------------------- a5: src/multinomial.cpp:13-32 incl. std::random_shuffle
@pytest.mark.parametrize("N", [1, 2, 33, 256, 777, 4096])
def test_multinomial_equals_reference(oracle, ref, N):
    rng = np.random.default_rng(100 + N)
    for kind in KINDS:
        w = weights(rng, N, kind)
        for n in (N, max(1, N // 2), 2 * N):
            ur = rng.random(n); us = rng.random(max(n - 1, 0))
            a_ref, oob = ref.multinomial_resampling_n_(w, n, ur, us)
            assert oob == 0
            a_orc = oracle.multinomial_resampling_n_(w, n, ur, us if n > 1 else None)
            assert a_ref.max() < N
            assert np.array_equal(a_orc, a_ref), (kind, n)


def test_multinomial_unnormalised_weights_equal_reference(oracle, ref):
    """:19,24 scale the sorted uniforms by sum(w): the reference accepts any positive weights."""
    rng = np.random.default_rng(5)
    for N in (7, 300):
        w = rng.random(N) * 37.5
        ur = rng.random(N); us = rng.random(N - 1)
        a_ref, _ = ref.multinomial_resampling_n_(w, N, ur, us)
        assert np.array_equal(oracle.multinomial_resampling_n_(w, N, ur, us), a_ref)


# ---------------------------------------------------------------- a6: src/multinomial.cpp:35-71
@pytest.mark.parametrize("N", [1, 2, 33, 256, 1000])
def test_coupled_multinomial_equals_reference(oracle, ref, N):
    rng = np.random.default_rng(200 + N)
    for k1 in KINDS:
        for k2 in ("1", "onehot", "zeros"):
            w1, w2 = weights(rng, N, k1), weights(rng, N, k2)
            for n in (N, max(1, N // 2)):
                ur = rng.random(n); us = rng.random(max(n - 1, 0))
                a_ref, oob = ref.coupled_multinomial_resampling_n_(w1, w2, n, ur, us)
                a_orc = oracle.coupled_multinomial_resampling_n_(w1, w2, n, ur, us if n > 1 else None)
                # not scaled by sum(w) (:49): with sum(w) < e_(n) the reference returns N (one past the end); the oracle clamps (H8)
                assert np.array_equal(np.minimum(a_ref, N - 1), a_orc), (k1, k2, n)


# ---------------------------------------------------------------- a7: src/coupledresamplingindexmatching.cpp:8-57
@pytest.mark.parametrize("N", [2, 33, 256, 1000])
def test_indexmatching_equals_reference(oracle, ref, N):
    rng = np.random.default_rng(300 + N)
    for k1 in ("0.1", "1", "5", "zeros", "uniform"):
        for k2 in ("0.1", "1", "zeros"):
            w1, w2 = weights(rng, N, k1), weights(rng, N, k2)
            for n in (N, 1, max(1, N // 2)):
                uc = rng.random(N)
                alpha = 0.0
                for a, b in zip(w1, w2):
                    alpha += min(a, b)
                nc = int(np.sum(uc[:n] < alpha)); nres = n - nc
                um, ums = rng.random(nc), rng.random(max(nc - 1, 0))
                ur, urs = rng.random(nres), rng.random(max(nres - 1, 0))
                stream = np.concatenate([uc, um, ums, ur, urs])
                a_ref, oob = ref.indexmatching(w1, w2, n, stream)
                a_orc = oracle.indexmatching(w1, w2, n, uc, um if nc else None, ums if nc > 1 else None, ur if nres else None, urs if nres > 1 else None)
                assert np.array_equal(np.minimum(a_ref, N - 1), a_orc), (k1, k2, n)


def test_indexmatching_identical_weights_equal_reference(oracle, ref):
    """alpha = 1 up to rounding: 1 - alpha is 0 or ~1e-16 and the residuals are 0/0 or garbage; every draw is coupled unless a
    uniform falls in [alpha, 1).  Oracle and reference must agree wherever the reference stays inside its arrays."""
    rng = np.random.default_rng(9)
    for N in (16, 500):
        w = weights(rng, N, "1")
        uc = rng.random(N) * 0.999
        um, ums = rng.random(N), rng.random(N - 1)
        a_ref, oob = ref.indexmatching(w, w, N, np.concatenate([uc, um, ums]))
        assert oob == 0
        a_orc = oracle.indexmatching(w, w, N, uc, um, ums, None, None)
        assert np.array_equal(a_ref, a_orc) and np.array_equal(a_ref[:, 0], a_ref[:, 1])


# ---------------------------------------------------------------- a9: src/tree.cpp:6-141
@pytest.mark.parametrize("N,M,d", [(8, 0, 1), (64, 0, 2), (100, 300, 3), (256, 2560, 1), (1000, 0, 4)])
def test_tree_slot_arrays_equal_reference(oracle, ref, N, M, d):
    """40 updates with multinomial-like ancestors; M0 = 3N forces Tree::double_size (:54-61, :101-116) for the degenerate and
    the near-identity ancestor patterns.  a_star / o_star / l_star / x_star and every path must be identical."""
    rng = np.random.default_rng(N + d)
    to, tr = oracle.Tree(N, M, d), ref.Tree(N, M, d)
    x0 = rng.normal(size=(d, N))
    to.init(x0); tr.init(x0)
    M0 = tr.M
    for step in range(40):
        if step % 10 == 7:
            a = np.arange(N, dtype=np.int32)                      # no coalescence: the tree must grow
        elif step % 10 == 3:
            a = np.full(N, rng.integers(N), dtype=np.int32)       # one survivor
        else:
            w = np.exp(rng.normal(size=N) * (0.3 if step % 2 else 2.0)); w /= w.sum()
            a = np.sort(rng.choice(N, size=N, p=w)).astype(np.int32)
            if step % 3 == 0:
                rng.shuffle(a)
        x = rng.normal(size=(d, N))
        to.update(x, a); tr.update(x, a)
        ar, orr, lr, xr = tr.arrays()
        assert to.M == tr.M and to.nsteps == tr.nsteps == step + 1
        assert np.array_equal(to.l_star, lr), step
        live = np.zeros(tr.M, bool)                                # compare the nodes reachable from the leaves ...
        for j in lr:
            while j >= 0 and not live[j]:
                live[j] = True; j = ar[j]
        assert np.array_equal(to.a_star[live], ar[live]) and np.array_equal(to.o_star[live], orr[live]), step
        assert np.array_equal(to.a_star, ar) and np.array_equal(to.o_star, orr), step   # ... and the dead slots too
    for n in (0, N // 2, N - 1):
        assert np.array_equal(to.get_path(n), tr.get_path(n))
    for lag in (0, 1, 5):
        assert np.array_equal(to.retrieve_xgeneration(lag), tr.retrieve_xgeneration(lag))


def test_tree_double_size_equals_reference(oracle, ref):
    """No coalescence at all (identity ancestors): nothing is pruned, 3N slots fill after two updates and Tree::double_size
    (:54-61, :101-116) runs, twice over 8 updates."""
    N, d = 50, 2
    rng = np.random.default_rng(0)
    to, tr = oracle.Tree(N, 0, d), ref.Tree(N, 0, d)
    x0 = rng.normal(size=(d, N)); to.init(x0); tr.init(x0)
    a = np.arange(N, dtype=np.int32)
    for step in range(8):
        x = rng.normal(size=(d, N))
        to.update(x, a); tr.update(x, a)
        ar, orr, lr, xr = tr.arrays()
        assert to.M == tr.M and np.array_equal(to.l_star, lr) and np.array_equal(to.a_star, ar) and np.array_equal(to.o_star, orr)
    assert tr.M == 12 * N
    for n in (0, 17, N - 1):
        assert np.array_equal(to.get_path(n), tr.get_path(n))


# ---------------------------------------------------------------- a12 pieces that compile: create_A_, SVLevy
@pytest.mark.parametrize("d", [1, 2, 5, 10, 16])
def test_create_A_equals_reference(oracle, ref, d):
    for alpha in (0.5, 0.4, 0.95):
        assert np.array_equal(oracle.create_A(alpha, d), ref.create_A(alpha, d))


def _sv_cfg(oracle, N, theta):
    return oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, N, 5, theta=theta)


def test_levy_sv_transition_equals_reference(oracle, ref):
    """src/SVLevy_functions.cpp:58-70 with the Poisson counts, jump times and jump sizes injected.  The oracle takes the two
    per-particle sums (sum exp(-lambda (t - c)) e, sum e) as its injected noise (DESIGN.md "noise injection"): they are formed
    here exactly as Rcpp sugar does (left-to-right double sums of the elementwise products)."""
    import math
    rng = np.random.default_rng(11)
    theta = [0.244872, -0.275641, 0.816154, 0.093333, 0.051538]
    mu, beta, xi, w2, lam = theta
    N = 500
    X = np.stack([rng.gamma(2.0, 0.3, N), rng.gamma(2.0, 0.3, N)], axis=1)       # (N, 2) = (v, z)
    k = rng.poisson(1.3, N).astype(float)
    nj = int(k.sum())
    cu, eu = rng.random(nj), rng.exponential(size=nj)
    t_1, t = 3.0, 4.0
    Xn = ref.rtransition_SVLevy(X, t_1, t, xi, w2, lam, k, cu, eu)
    swe, se = np.zeros(N), np.zeros(N)
    pos = 0
    scale = 1.0 / (xi / w2)
    for i in range(N):
        a = b = 0.0
        for j in range(int(k[i])):
            c = t_1 + (t - t_1) * cu[pos]; e = scale * eu[pos]; pos += 1
            a += math.exp(-lam * (t - c)) * e
            b += e
        swe[i], se[i] = a, b
    cfg = _sv_cfg(oracle, N, theta)
    xo = oracle.model_rtransition(cfg, X.T.copy(), 4, np.stack([swe, se]))       # oracle state is (d, N)
    assert np.array_equal(xo.T, Xn)
    # rinitial (:40-55): z0 ~ Gamma injected, same sums with t_1 = 0
    z0 = rng.gamma(xi * xi / w2, w2 / xi, N)
    Xi = ref.rinitial_SVLevy(N, 1.0, xi, w2, lam, z0, k, cu, eu)
    pos = 0
    for i in range(N):
        a = b = 0.0
        for j in range(int(k[i])):
            c = 0.0 + (1.0 - 0.0) * cu[pos]; e = scale * eu[pos]; pos += 1
            a += math.exp(-lam * (1.0 - c)) * e
            b += e
        swe[i], se[i] = a, b
    xo = oracle.model_rinit(cfg, np.stack([z0, swe, se]))
    assert np.array_equal(xo.T, Xi)


def test_levy_sv_measurement_equals_reference(oracle, ref):
    """src/SVLevy_functions.cpp:73-81 (R's dnorm formula is the stub's restatement: 1e-15, not bit, is the claim here)."""
    rng = np.random.default_rng(12)
    theta = [0.244872, -0.275641, 0.816154, 0.093333, 0.051538]
    N = 300
    X = np.stack([rng.gamma(2.0, 0.3, N), rng.gamma(2.0, 0.3, N)], axis=1)
    cfg = _sv_cfg(oracle, N, theta)
    for y in (0.3, -2.5):
        lw_ref = ref.dobs_SVLevy(y, X, theta[0], theta[1], True)
        lw_orc = oracle.model_dmeasurement(cfg, X.T.copy(), [y])
        assert np.max(np.abs(lw_ref - lw_orc) / np.abs(lw_ref)) < 1e-14
