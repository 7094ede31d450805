"""GPU parity of the wide-state filter (Lorenz 96, d = 32 / 64, one filter across the GPU; BASELINE config 5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def up():
    import unbiased_pimh_b200 as m
    return m


@pytest.mark.parametrize("d,N,substeps", [(64, 96, 1), (64, 200, 3), (32, 130, 2)])
def test_wide_injected_noise_parity(up, oracle, d, N, substeps):
    rng = np.random.default_rng(d + N)
    T, B = 12, 2
    y = rng.normal(size=(T, d)) * 2 + 2
    y[5, :] = np.nan                                            # NA row: weights 0, and no resampling before step 7
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, T, theta=[8.0, 0.4, 1.2], integrator=0, substeps=substeps)
    nz = dict(init=rng.normal(size=(B, d, N)), trans=rng.normal(size=(B, T, d, N)), resample=np.sort(rng.random((B, T, N)), axis=2), path_u=rng.random(B))
    pb = up.PfBatch(up.get_lorenz(1.2, d, "rk4", substeps), [8.0, 0.4], y, N, B)
    pb.record_ancestors(True)
    pb.set_noise(B, nz)
    pb.run(B, 0)
    r = pb.fetch(logz_t=True, path_index=True)
    for b in range(B):
        o = oracle.cpf(ocfg, y, noise={k: v[b] for k, v in nz.items()}, want_ancestors=True)
        assert np.array_equal(pb.fetch_ancestors(b), o["ancestors"])
        assert np.allclose(r["logz_t"][b], o["logz_t"], rtol=1e-10, atol=1e-300)          # 1e-10 relative, fp64
        assert r["logz_t"][b][5] == 0.0
        assert int(r["path_index"][b]) == o["path_index"]
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)
    pb.close()


@pytest.mark.parametrize("d,N", [(64, 300), (64, 5000), (32, 257)])
def test_wide_philox_parity(up, oracle, d, N):
    rng = np.random.default_rng(d * 7 + N)
    T, B, seed = 10, 2, 99
    y = rng.normal(size=(T, d)) * 2 + 2
    r = up.cpf_batch(N, up.get_lorenz(1.0, d, "rk4", 2), [8.0, 0.5], y, B, seed=seed, first_filter=3, logz_t=True)
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, T, theta=[8.0, 0.5, 1.0], integrator=0, substeps=2)
    for b in range(B):
        o = oracle.cpf(ocfg, y, seed=seed, filter_id=3 + b)
        assert np.allclose(r["logz_t"][b], o["logz_t"], rtol=1e-10, atol=0)
        assert int(r["path_index"][b]) == o["path_index"]
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)


def test_wide_full_size_properties(up):
    """BASELINE configs[4] shape (d = 64, N = 65536), short horizon: ancestors sorted and in range, log Z finite,
    bit-reproducible, the drawn path is a row of the stored history."""
    rng = np.random.default_rng(1)
    d, N, T = 64, 65536, 8
    y = rng.normal(size=(T, d)) * 1.5 + 2
    pb = up.PfBatch(up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, N, 1)
    pb.record_ancestors(True)
    pb.run(1, 7)
    a = pb.fetch(logz_t=True, path_index=True)
    anc = pb.fetch_ancestors(0)
    assert np.array_equal(anc[0], np.arange(N)) and anc.min() >= 0 and anc.max() < N and np.all(np.diff(anc[1:], axis=1) >= 0)
    assert np.all(np.isfinite(a["logz_t"])) and np.isfinite(a["logz"][0])
    pb.run(1, 7)
    b = pb.fetch(logz_t=True, path_index=True)
    assert np.array_equal(a["logz_t"], b["logz_t"]) and np.array_equal(a["trajectory"], b["trajectory"])
    assert a["trajectory"].shape == (1, d, T + 1) and np.all(np.abs(a["trajectory"]) < 50)
    pb.close()


def test_wide_philox_parity_with_missing_observation(up, oracle):
    """The Philox (two-level CDF) path through an NA row: weights 0 at that step, no resampling at the next one
    (R/CPF.R:38-40, R/model_get_lorenz.R:39-40) -- the launch after the NA step must still write log Z of the NA step."""
    d, N, T, seed = 64, 5000, 8, 5
    y = np.random.default_rng(3).normal(size=(T, d)) * 2 + 2
    y[3, :] = np.nan
    r = up.cpf_batch(N, up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, 1, seed=seed, logz_t=True)
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, T, theta=[8.0, 0.5, 1.0], integrator=0, substeps=1)
    o = oracle.cpf(ocfg, y, seed=seed, filter_id=0)
    assert r["logz_t"][0][3] == 0.0
    assert np.allclose(r["logz_t"][0], o["logz_t"], rtol=1e-10, atol=1e-300)
    assert int(r["path_index"][0]) == o["path_index"]
    assert np.allclose(r["trajectory"][0], o["trajectory"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("N,B,G", [(20000, 3, 0), (1000, 1, 5), (1000, 2, 32), (4097, 1, 14)])
def test_wide_group_sizes(up, oracle, monkeypatch, N, B, G):
    """Any group size 1..32 (not only powers of two; the last group ragged): G = 0 takes the library's own choice
    (13 for 3 x 20000 particles on 148 SMs), the others force it through the tuning knob."""
    if G:
        monkeypatch.setenv("CPIMH_WIDE_G", str(G))
    d, T, seed = 64, 5, 11
    y = np.random.default_rng(N).normal(size=(T, d)) * 2 + 2
    pb = up.PfBatch(up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, N, B)
    eff = pb.effective_config()
    if G:
        assert eff["threads_per_filter"] == G                      # wide filter: particles per warp group
    else:
        assert eff["threads_per_filter"] not in (1, 2, 4, 8, 16, 32)
    pb.run(B, seed, first_filter=2)
    r = pb.fetch(logz_t=True, path_index=True)
    pb.close()
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, T, theta=[8.0, 0.5, 1.0], integrator=0, substeps=1)
    for b in range(B):
        o = oracle.cpf(ocfg, y, seed=seed, filter_id=2 + b)
        assert np.allclose(r["logz_t"][b], o["logz_t"], rtol=1e-10, atol=0)
        assert int(r["path_index"][b]) == o["path_index"]
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)


def test_wide_more_groups_than_the_shared_memory_table_holds(up):
    """N = 300 000 gives 9375 groups of 32 > the 8192 a CTA's shared memory holds: the per-particle cluster scan runs
    between the steps instead.  Properties only (ancestors sorted and in range, finite log Z, reproducible)."""
    d, N, T = 64, 300000, 4
    y = np.random.default_rng(2).normal(size=(T, d)) * 1.5 + 2
    pb = up.PfBatch(up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, N, 1)
    assert pb.effective_config()["ctas_per_sm"] > 8192             # wide filter: number of warp groups
    pb.record_ancestors(True)
    pb.run(1, 7)
    a = pb.fetch(logz_t=True, path_index=True)
    anc = pb.fetch_ancestors(0)
    assert np.array_equal(anc[0], np.arange(N)) and anc.min() >= 0 and anc.max() < N and np.all(np.diff(anc[1:], axis=1) >= 0)
    assert np.all(np.isfinite(a["logz_t"]))
    pb.run(1, 7)
    b = pb.fetch(logz_t=True, path_index=True)
    assert np.array_equal(a["logz_t"], b["logz_t"]) and np.array_equal(a["trajectory"], b["trajectory"])
    pb.close()


def _wide_run(up, y, N, B, seed, noise=None, d=64, exact=False):
    pb = up.PfBatch(up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, N, B)
    pb.record_ancestors(True)
    if exact:
        pb.exact_mode(True)
    if noise is not None:
        pb.set_noise(B, noise)
    pb.run(B, seed, first_filter=3)
    r = pb.fetch(logz_t=True, path_index=True)
    r["anc"] = np.stack([pb.fetch_ancestors(b) for b in range(B)])
    r["rep"] = pb.fetch_exact_repeats(B)
    pb.close()
    return r


@pytest.mark.parametrize("N,B", [(5000, 2), (65536, 1)])
def test_wide_reference_order_mode_equals_the_fast_pass(up, N, B):
    """Exact ancestors, wide filter: the fast pass (group-table CDF, parallel sums) compares every chosen knot with the uniform
    +- the worst-case rounding margin and reports per filter whether any draw was that close; cpimh_pf_exact_mode(h, 1) runs
    every step with cumsum(normweights) in the reference's left-to-right order (src/multinomial.cpp:17) instead.  Both modes must
    give the same ancestors, path and trajectory on random draws (log Z: another summation order, 1e-13); at N = 5000 no draw
    comes within the margin (2 N^2 m ~ 3e-4 per step), at N = 65536 about 0.7 per step do -- which is why the mode is opt-in."""
    T = 6 if N < 10000 else 4
    y = np.random.default_rng(N).normal(size=(T, 64)) * 2 + 2
    fast = _wide_run(up, y, N, B, 19)
    if N < 10000:
        assert not fast["rep"].any()
    ref = _wide_run(up, y, N, B, 19, exact=True)
    assert not ref["rep"].any()
    assert np.array_equal(fast["anc"], ref["anc"])
    assert np.array_equal(fast["path_index"], ref["path_index"]) and np.array_equal(fast["trajectory"], ref["trajectory"])
    assert np.allclose(fast["logz_t"], ref["logz_t"], rtol=1e-13, atol=0)


def test_wide_uniform_on_a_cdf_knot_is_flagged_and_settled_by_the_reference_order_mode(up, oracle, monkeypatch):
    """Sorted uniforms placed ON (or one ulp beside) knots of the step-1 CDF: the fast pass must flag the filter; the
    reference-order mode (API switch or CPIMH_WIDE_EXACT=1) settles those draws with the left-to-right CDF.  (The knots come from
    the oracle's weights, equal to the kernel's to ~1e-15: far inside the margin, so the check trips; the kernel's own last bits
    decide the crafted ancestors, every other draw must agree with the oracle.)"""
    from tests.test_gpu_exact_ancestors import craft_sorted_uniforms
    d, N, B = 64, 2000, 2
    rng = np.random.default_rng(8)
    y = rng.normal(size=(2, d)) * 2 + 2
    nz = dict(init=rng.normal(size=(B, d, N)), trans=rng.normal(size=(B, 2, d, N)), resample=np.zeros((B, 2, N)), path_u=rng.random(B))
    nz["resample"][:, 0, :] = 0.5
    ocfg1 = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, 1, theta=[8.0, 0.5, 1.0], integrator=0, substeps=1)
    ncraft = 40
    for b in range(B):
        n1 = {"init": nz["init"][b], "trans": nz["trans"][b][:1], "resample": nz["resample"][b][:1], "path_u": nz["path_u"][b]}
        w = oracle.cpf(ocfg1, y[:1], noise=n1, all_particles=True)["normweights"]
        e, hits = craft_sorted_uniforms(rng, w, ncraft)
        assert hits >= 10
        nz["resample"][b, 1] = e
    fast = _wide_run(up, y, N, B, 0, noise=nz)
    assert fast["rep"].all(), "a uniform on a knot must be reported"
    ref = _wide_run(up, y, N, B, 0, noise=nz, exact=True)
    assert not ref["rep"].any()
    monkeypatch.setenv("CPIMH_WIDE_EXACT", "1")
    ref2 = _wide_run(up, y, N, B, 0, noise=nz)
    assert np.array_equal(ref["anc"], ref2["anc"]) and np.array_equal(ref["trajectory"], ref2["trajectory"]) and np.array_equal(ref["logz_t"], ref2["logz_t"])
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, 2, theta=[8.0, 0.5, 1.0], integrator=0, substeps=1)
    for b in range(B):
        o = oracle.cpf(ocfg, y, noise={k: v[b] for k, v in nz.items()}, want_ancestors=True)
        assert int(np.sum(ref["anc"][b][1] != o["ancestors"][1])) <= ncraft        # only crafted draws may land on the other side
        assert int(np.sum(fast["anc"][b][1] != ref["anc"][b][1])) <= ncraft


@pytest.mark.parametrize("d,N", [(64, 300), (32, 1000)])
def test_wide_all_particle_average_matches_oracle(up, oracle, d, N):
    """all_particles = TRUE (R/CPF.R:73-76) for the wide filter: the weighted average of all N paths, walked from the leaves."""
    T, B, seed = 9, 2, 13
    y = np.random.default_rng(d + N).normal(size=(T, d)) * 2 + 2
    r = up.cpf_batch(N, up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, B, seed=seed, first_filter=2, all_particles=True)
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, d, d, N, T, theta=[8.0, 0.5, 1.0], integrator=0, substeps=1)
    for b in range(B):
        o = oracle.cpf(ocfg, y, seed=seed, filter_id=2 + b, all_particles=True)
        assert np.allclose(r["exp_path"][b], o["exp_path"], rtol=1e-9, atol=1e-9)
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)
    again = up.cpf_batch(N, up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5], y, B, seed=seed, first_filter=2, all_particles=True)
    assert np.array_equal(again["exp_path"], r["exp_path"])                         # fixed summation order


def test_r_shaped_pimh_chains_on_the_wide_filter(up):
    """get_pimh_kernels / coupled_pimh_chains (R/pf_kernels.R:88-145, R/debiasing_algorithm.R:140-265) with Lorenz 96 d = 64:
    every proposal is one wide filter; also with all_particles = TRUE (the Rao-Blackwellised state)."""
    d, N, T = 64, 2000, 4
    y = np.random.default_rng(21).normal(size=(T, d)) * 1.5 + 2
    model, th = up.get_lorenz(1.0, d, "rk4", 1), [8.0, 0.5]
    up.set_seed(11)
    ker = up.get_pimh_kernels(False, model=model, theta=th, observations=y)
    c = up.coupled_pimh_chains(ker["single_kernel"], ker["coupled_kernel"], ker["pimh_init"], N, K=2, max_iterations=200)
    assert c["finished"] and c["meetingtime"] >= 1 and c["samples1"].shape[1] == T + 1
    hb = up.H_bar(c, k=1, K=2)
    assert np.all(np.isfinite(hb))
    kap = up.get_pimh_kernels(False, all_particles=True, model=model, theta=th, observations=y)
    c2 = up.coupled_pimh_chains(kap["single_kernel"], kap["coupled_kernel"], kap["pimh_init"], N, K=2, max_iterations=200)
    assert c2["finished"] and c2["exp_samples1"].shape == c2["samples1"].shape and np.all(np.isfinite(c2["exp_samples1"]))
