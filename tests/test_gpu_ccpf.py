"""GPU tests of the conditional particle filters: CPF(ref_trajectory, with_as), CPF_coupled with index-matching
resampling, and coupled_ccpf_chains -- compared with the oracle (oracle/ccpf.c) on the same Philox streams."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A_OBS = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
SK_THETA = {"A_obs": A_OBS, "s2_obs": 1.0}
TH_AR = [0.5, math.sqrt(10.0)]


@pytest.fixture(scope="module")
def up():
    import unbiased_pimh_b200 as m
    return m


def _refs(oracle, cfg, y, seed, n):
    return [oracle.cpf(cfg, y, seed=seed, filter_id=100 + i)["trajectory"] for i in range(n)]


@pytest.mark.parametrize("name,with_as", [("ar1", False), ("ar1", True), ("gss", False), ("gss", True), ("ar3", True), ("sv", False), ("sk", False)])
def test_conditional_filter_matches_oracle(up, oracle, golden, name, with_as):
    """One system: same streams => same ancestors (the last one pinned to N-1 or ancestor-sampled), same trajectory and log Z."""
    if name == "ar1":
        m, th, y, N, oc = up.get_ar(1), TH_AR, golden["lgssm_y"], 64, (oracle.MODEL_AR, 1, 1)
    elif name == "ar3":
        m, th, y, N, oc = up.get_ar(3), TH_AR, np.random.default_rng(4).normal(size=(30, 3)), 100, (oracle.MODEL_AR, 3, 3)
    elif name == "gss":
        m, th, y, N, oc = up.get_gss(), [], golden["gss_y"][:60], 128, (oracle.MODEL_GSS, 1, 1)
    elif name == "sv":
        m, th, y, N, oc = up.get_levy_sv(), golden["sv_theta_mle"], golden["sv_y"][:40], 96, (oracle.MODEL_LEVY_SV, 2, 1)
    else:
        m, th, y, N, oc = up.get_stochastic_kinetic(), SK_THETA, golden["sk_y"][:20], 64, (oracle.MODEL_SK, 4, 2)
    T = y.shape[0]
    oth = th if name != "sk" else [1.0] + list(A_OBS.flatten(order="F"))
    cfg = oracle.make_config(oc[0], oc[1], oc[2], N, T, theta=oth)
    refs = _refs(oracle, cfg, y, 9, 3)
    g = up.ccpf_batch(N, m, th, y, refs, with_as=with_as, seed=21, first_filter=4, ancestors=True)
    for b in range(3):
        o = oracle.ccpf(cfg, y, refs[b], with_as=with_as, seed=21, filter_id=4 + b, want_ancestors=True)
        assert np.array_equal(g["ancestors"][b, 0], o["ancestors1"])
        assert g["path_index"][b, 0] == o["path_index"][0]
        assert np.allclose(g["trajectory1"][b], o["trajectory1"], rtol=1e-9, atol=1e-9)
        assert abs(g["logz"][b] - o["logz"]) <= 1e-9 * abs(o["logz"])
        if not with_as:
            assert np.all(g["ancestors"][b, 0][:, -1] == N - 1)


@pytest.mark.parametrize("name,with_as", [("ar1", False), ("ar1", True), ("gss", True), ("ar3", False), ("sv", False), ("sk", False)])
def test_coupled_conditional_filter_matches_oracle(up, oracle, golden, name, with_as):
    """Two systems, index-matching resampling, common random numbers: both ancestor tables and both trajectories equal the oracle's."""
    if name == "ar1":
        m, th, y, N, oc = up.get_ar(1), TH_AR, golden["lgssm_y"], 64, (oracle.MODEL_AR, 1, 1)
    elif name == "ar3":
        m, th, y, N, oc = up.get_ar(3), TH_AR, np.random.default_rng(4).normal(size=(30, 3)), 100, (oracle.MODEL_AR, 3, 3)
    elif name == "gss":
        m, th, y, N, oc = up.get_gss(), [], golden["gss_y"][:60], 128, (oracle.MODEL_GSS, 1, 1)
    elif name == "sv":
        m, th, y, N, oc = up.get_levy_sv(), golden["sv_theta_mle"], golden["sv_y"][:40], 96, (oracle.MODEL_LEVY_SV, 2, 1)
    else:
        m, th, y, N, oc = up.get_stochastic_kinetic(), SK_THETA, golden["sk_y"][:20], 64, (oracle.MODEL_SK, 4, 2)
    T = y.shape[0]
    oth = th if name != "sk" else [1.0] + list(A_OBS.flatten(order="F"))
    cfg = oracle.make_config(oc[0], oc[1], oc[2], N, T, theta=oth)
    refs = _refs(oracle, cfg, y, 9, 6)
    g = up.ccpf_batch(N, m, th, y, refs[:3], refs[3:], with_as=with_as, seed=22, first_filter=1, ancestors=True)
    for b in range(3):
        o = oracle.ccpf(cfg, y, refs[b], refs[3 + b], with_as=with_as, seed=22, filter_id=1 + b, want_ancestors=True)
        assert np.array_equal(g["ancestors"][b, 0], o["ancestors1"]) and np.array_equal(g["ancestors"][b, 1], o["ancestors2"])
        assert list(g["path_index"][b]) == o["path_index"]
        assert np.allclose(g["trajectory1"][b], o["trajectory1"], rtol=1e-9, atol=1e-9)
        assert np.allclose(g["trajectory2"][b], o["trajectory2"], rtol=1e-9, atol=1e-9)


def test_injected_noise_and_faithful_coupling(up, oracle, golden):
    """Injected init/transition noise (the reference's rinit_rand / rtransition_rand arrays) gives the oracle's result, and two
    systems started from the SAME reference return bitwise identical trajectories (the chains stay together once met)."""
    y = golden["lgssm_y"][:50]
    N, T, B = 80, 50, 2
    rng = np.random.default_rng(3)
    noise = {"init": rng.normal(size=(B, 1, N)), "trans": rng.normal(size=(B, T, 1, N))}
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, N, T, theta=TH_AR)
    refs = _refs(oracle, cfg, y, 5, 4)
    g = up.ccpf_batch(N, up.get_ar(1), TH_AR, y, refs[:2], refs[2:], with_as=True, seed=6, noise=noise, ancestors=True)
    for b in range(B):
        o = oracle.ccpf(cfg, y, refs[b], refs[2 + b], with_as=True, seed=6, filter_id=b, noise={"init": noise["init"][b], "trans": noise["trans"][b]},
                        want_ancestors=True)
        assert np.array_equal(g["ancestors"][b, 0], o["ancestors1"]) and np.array_equal(g["ancestors"][b, 1], o["ancestors2"])
        assert np.allclose(g["trajectory1"][b], o["trajectory1"], rtol=1e-10, atol=1e-10)
        assert np.allclose(g["trajectory2"][b], o["trajectory2"], rtol=1e-10, atol=1e-10)
    same = up.ccpf_batch(N, up.get_ar(1), TH_AR, y, refs[:2], refs[:2], with_as=False, seed=8)
    assert np.array_equal(same["trajectory1"], same["trajectory2"])


def test_r_shaped_conditional_api_and_errors(up, golden):
    y = golden["lgssm_y"][:30]
    m = up.get_ar(1)
    up.set_seed(5)
    x0 = up.CPF(50, m, TH_AR, y)["trajectory"]
    r = up.CPF(50, m, TH_AR, y, ref_trajectory=x0, with_as=True)
    assert r["trajectory"].shape == (1, 31) and np.isfinite(r["logz"])
    c = up.CPF_coupled(50, m, TH_AR, y, x0, r["trajectory"], up.CR_indexmatching, with_as=False)
    assert c["new_trajectory1"].shape == (1, 31) and c["new_trajectory2"].shape == (1, 31)
    ker = up.get_cpf_kernels(True, model=m, theta=TH_AR, observations=y)
    init = up.get_pimh_kernels(False, model=m, theta=TH_AR, observations=y)["pimh_init"]
    ch = up.coupled_ccpf_chains(ker["single_kernel"], ker["coupled_kernel"], init, 50, K=3, max_iterations=200)
    assert ch["finished"] and ch["samples1"].shape == (ch["iteration"] + 1, 31) and ch["samples2"].shape == (ch["iteration"], 31)
    tau = ch["meetingtime"]
    for t in range(int(tau), ch["iteration"] + 1):
        assert np.array_equal(ch["samples1"][t], ch["samples2"][t - 1])
    assert up.H_bar(ch, k=1, K=3).shape == (31,)
    with pytest.raises(up.CpimhError):
        up.ccpf_batch(50, m, TH_AR, y, [x0], shuffle=1)
    with pytest.raises(up.CpimhError):
        ysv = golden["sv_y"][:10]
        xs = up.CPF(64, up.get_levy_sv(), golden["sv_theta_mle"], ysv)["trajectory"]
        up.CPF(64, up.get_levy_sv(), golden["sv_theta_mle"], ysv, ref_trajectory=xs, with_as=True)


@pytest.mark.parametrize("with_as", [False, True])
def test_ccpf_chains_match_oracle(up, oracle, golden, with_as):
    y = golden["lgssm_y"]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=TH_AR)
    for (k, K) in ((0, 1), (2, 5)):
        g = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, 64, 40, K=K, k=k, with_as=with_as, seed=12, first_replicate=3, max_iterations=400)
        o = oracle.coupled_ccpf_batch(cfg, y, 12, 3, 40, k=k, K=K, with_as=with_as, max_iterations=400)
        assert np.array_equal(g["meetingtime"], o["meetingtime"]) and np.array_equal(g["iteration"], o["iteration"])
        assert g["nfilters"] == o["nfilters"]
        assert np.allclose(g["hbar"][:, 0, :], o["hbar"], rtol=1e-9, atol=1e-9)
        assert np.allclose(g["sum_hbar"][0], o["hbar"].sum(axis=0), rtol=1e-9, atol=1e-8)


def test_ccpf_estimator_is_unbiased_for_the_smoothing_mean(up, golden):
    """H_bar averaged over replicates matches the Kalman/RTS smoothing mean of the LGSSM (x0 ~ N(0,1), a = 0.5, R = 10)."""
    y = golden["lgssm_y"]
    T, a, Rv = 100, 0.5, 10.0
    mp, Pp, mf, Pf = np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1)
    mf[0], Pf[0] = 0.0, 1.0
    for t in range(1, T + 1):
        mp[t], Pp[t] = a * mf[t - 1], a * a * Pf[t - 1] + 1
        Kg = Pp[t] / (Pp[t] + Rv)
        mf[t], Pf[t] = mp[t] + Kg * (y[t - 1, 0] - mp[t]), (1 - Kg) * Pp[t]
    ms = mf.copy()
    for t in range(T - 1, -1, -1):
        ms[t] = mf[t] + Pf[t] * a / Pp[t + 1] * (ms[t + 1] - mp[t + 1])
    R = 3000
    g = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, 128, R, K=8, k=3, with_as=True, seed=77)
    assert g["meetingtime"].min() >= 2 and g["meetingtime"].max() < 60
    est = g["hbar"][:, 0, :].mean(axis=0)
    se = g["hbar"][:, 0, :].std(axis=0, ddof=1) / math.sqrt(R)
    z = (est - ms) / se
    assert np.max(np.abs(z)) < 4.5 and np.mean(np.abs(z) > 2) < 0.15
    assert np.allclose(g["sum_hbar"][0] / R, est, rtol=1e-10, atol=1e-12)


def test_rao_blackwellised_conditional_filters_and_rheeglynn(up, oracle, golden):
    """CPF_RB / CPF_RB_coupled estimates (identity test function) equal the oracle's weighted average over all N paths; the chain
    driver's hbar_rb equals H_bar over the oracle's exp_samples; rheeglynn_estimator(_RB) are the k = K = 0 instances."""
    y = golden["lgssm_y"]
    N = 64
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, N, 100, theta=TH_AR)
    refs = _refs(oracle, cfg, y, 9, 4)
    g = up.ccpf_batch(N, up.get_ar(1), TH_AR, y, refs[:2], refs[2:], with_as=True, seed=30, exp_path=True)
    g1 = up.ccpf_batch(N, up.get_ar(1), TH_AR, y, refs[:2], with_as=False, seed=31, exp_path=True)
    for b in range(2):
        o = oracle.ccpf(cfg, y, refs[b], refs[2 + b], with_as=True, seed=30, filter_id=b, exp_path=True)
        assert np.allclose(g["exp_path1"][b], o["exp_path1"], rtol=1e-9, atol=1e-10)
        assert np.allclose(g["exp_path2"][b], o["exp_path2"], rtol=1e-9, atol=1e-10)
        o1 = oracle.ccpf(cfg, y, refs[b], with_as=False, seed=31, filter_id=b, exp_path=True)
        assert np.allclose(g1["exp_path1"][b], o1["exp_path1"], rtol=1e-9, atol=1e-10)
    for (k, K) in ((0, 0), (2, 2), (1, 4)):
        gg = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, N, 30, K=K, k=k, with_as=True, seed=12, first_replicate=3, rb=True)
        oo = oracle.coupled_ccpf_batch(cfg, y, 12, 3, 30, k=k, K=K, with_as=True, rb=True)
        assert np.array_equal(gg["meetingtime"], oo["meetingtime"])
        assert np.allclose(gg["hbar"][:, 0, :], oo["hbar"], rtol=1e-9, atol=1e-9)
        assert np.allclose(gg["hbar_rb"][:, 0, :], oo["hbar_rb"], rtol=1e-9, atol=1e-9)
        assert np.allclose(gg["sum_hbar_rb"][0], oo["hbar_rb"].sum(axis=0), rtol=1e-9, atol=1e-8)
    ap = {"nparticles": N, "with_as": True, "lambda": 0, "coupled_resampling": up.CR_indexmatching}
    e = up.rheeglynn_estimator(y, up.get_ar(1), TH_AR, ap, seed=12, replicate=3)
    erb = up.rheeglynn_estimator_RB(y, up.get_ar(1), TH_AR, ap, seed=12, replicate=3)
    o0 = oracle.coupled_ccpf_batch(cfg, y, 12, 3, 1, k=0, K=0, with_as=True, rb=True)
    assert np.allclose(e["estimate"][0], o0["hbar"][0], rtol=1e-9, atol=1e-9) and e["iteration"] == o0["iteration"][0]
    assert np.allclose(erb["estimate"][0], o0["hbar_rb"][0], rtol=1e-9, atol=1e-9)
    r = up.CPF_RB(N, up.get_ar(1), TH_AR, y, ref_trajectory=refs[0], with_as=True, seed=5)
    assert r["estimate"].shape == (1, 101) and r["new_trajectory"].shape == (1, 101)
    r0 = up.CPF_RB(N, up.get_ar(1), TH_AR, y, seed=5)
    assert r0["estimate"].shape == (1, 101)
    rc = up.CPF_RB_coupled(N, up.get_ar(1), TH_AR, y, refs[0], refs[1], with_as=False, seed=5)
    assert rc["estimate1"].shape == (1, 101) and not np.array_equal(rc["estimate1"], rc["estimate2"])


@pytest.mark.parametrize("name", ["ar1", "gss", "sv", "sk", "ar3"])
@pytest.mark.parametrize("ess,systematic", [(1.0, False), (0.5, False), (0.5, True)])
def test_adaptive_pf_matches_oracle(up, oracle, golden, name, ess, systematic):
    """pf() (R/CPF.R:106-155): same resampling decisions, ancestors, cumulative log-likelihood, filtering means and path."""
    if name == "ar1":
        m, th, y, N, oc = up.get_ar(1), TH_AR, golden["lgssm_y"], 96, (oracle.MODEL_AR, 1, 1)
    elif name == "ar3":
        m, th, y, N, oc = up.get_ar(3), TH_AR, np.random.default_rng(4).normal(size=(30, 3)), 100, (oracle.MODEL_AR, 3, 3)
    elif name == "gss":
        m, th, y, N, oc = up.get_gss(), [], golden["gss_y"][:60], 128, (oracle.MODEL_GSS, 1, 1)
    elif name == "sv":
        m, th, y, N, oc = up.get_levy_sv(), golden["sv_theta_mle"], golden["sv_y"][:40], 200, (oracle.MODEL_LEVY_SV, 2, 1)
    else:
        m, th, y, N, oc = up.get_stochastic_kinetic(), SK_THETA, golden["sk_y"][:20], 64, (oracle.MODEL_SK, 4, 2)
    T = y.shape[0]
    oth = th if name != "sk" else [1.0] + list(A_OBS.flatten(order="F"))
    cfg = oracle.make_config(oc[0], oc[1], oc[2], N, T, theta=oth)
    g = up.pf_batch(y, th, m, N, 3, ess_threshold=ess, systematic=systematic, seed=41, first_filter=2, ancestors=True)
    for b in range(3):
        o = oracle.pf_adaptive(cfg, y, seed=41, filter_id=2 + b, ess_threshold=ess, systematic=systematic, want_ancestors=True)
        assert np.array_equal(g["resampled"][b], o["resampled"])
        assert np.array_equal(g["ancestors"][b], o["ancestors"])
        assert np.allclose(g["loglik_t"][b], o["loglik_t"], rtol=1e-10, atol=1e-10) and abs(g["loglik"][b] - o["loglik"]) <= 1e-10 * abs(o["loglik"])
        assert np.allclose(g["pf_means"][b], o["pf_means"], rtol=1e-9, atol=1e-10)
        assert np.allclose(g["path"][b], o["path"], rtol=1e-9, atol=1e-9)
    if ess == 1.0:
        assert g["resampled"].all()
    else:
        assert 0 < g["resampled"].mean() < 1


@pytest.mark.parametrize("name,N", [("gss", 128), ("sv", 2000), ("ar3", 1000)])
@pytest.mark.parametrize("ess,systematic", [(1.0, False), (0.6, True)])
def test_adaptive_pf_reference_order_mode_equals_the_fast_pass(up, golden, monkeypatch, name, N, ess, systematic):
    """Exact ancestors in pf(): every chosen knot of the parallel-scan CDF is compared with the uniform +- a rounding margin; a
    step that trips the check is repeated inside the kernel with cumsum(normweights) left to right (src/multinomial.cpp:17,
    src/systematic.cpp:13-19).  On random draws nothing trips, and forcing the reference-order mode must return the same bits."""
    if name == "gss":
        m, th, y = up.get_gss(), [], golden["gss_y"][:40]
    elif name == "sv":
        m, th, y = up.get_levy_sv(), golden["sv_theta_mle"], golden["sv_y"][:25]
    else:
        m, th, y = up.get_ar(3), TH_AR, np.random.default_rng(4).normal(size=(30, 3))
    fast = up.pf_batch(y, th, m, N, 4, ess_threshold=ess, systematic=systematic, seed=5, first_filter=1, ancestors=True)
    assert up.lib().cpimh_last_exact_repeats() == 0
    # a margin widened 10^9-fold makes many steps look near a knot: each is repeated in reference order inside the kernel and
    # must come out with the same ancestors (the repeat restores the weights, redraws the same uniforms, sums left to right)
    monkeypatch.setenv("CPIMH_EXACT_MARGIN_SCALE", "1e9")
    rep = up.pf_batch(y, th, m, N, 4, ess_threshold=ess, systematic=systematic, seed=5, first_filter=1, ancestors=True)
    assert up.lib().cpimh_last_exact_repeats() > 0
    monkeypatch.delenv("CPIMH_EXACT_MARGIN_SCALE")
    monkeypatch.setenv("CPIMH_APF_EXACT", "1")
    ref = up.pf_batch(y, th, m, N, 4, ess_threshold=ess, systematic=systematic, seed=5, first_filter=1, ancestors=True)
    for k in ("ancestors", "resampled", "path", "loglik_t", "pf_means"):
        assert np.array_equal(fast[k], ref[k]), k
        assert np.array_equal(rep[k], ref[k]), k


def test_adaptive_pf_likelihood_is_unbiased(up, golden):
    """exp(loglik) is an unbiased estimator of the likelihood: log-mean-exp over 4000 filters matches the Kalman value."""
    y = golden["lgssm_y"]
    for ess, systematic in ((1.0, False), (0.5, True)):
        r = up.pf_batch(y, TH_AR, up.get_ar(1), 256, 4000, ess_threshold=ess, systematic=systematic, seed=9)
        ll = r["loglik"]
        lme = np.log(np.mean(np.exp(ll - ll.max()))) + ll.max()
        assert abs(lme - float(golden["lgssm_kalman_loglik"])) < 0.05
        assert np.allclose(r["loglik_t"][:, -1], ll, rtol=1e-13)
    one = up.pf(y, TH_AR, up.get_ar(1), 128, seed=3)
    assert one["pf_means"].shape == (100, 1) and one["path"].shape == (101, 1) and one["loglik_t"].shape == (100,)


def test_conditional_filters_with_missing_observations_and_lorenz(up, oracle):
    """R/CPF.R:38-40 / R/CPF_coupled.R:45-48: no resampling after a row whose first entry is NA (get_lorenz's dmeasurement
    returns 0 there); Lorenz-96 (d = 8, rk4) through the conditional kernels, one and two systems."""
    rng = np.random.default_rng(3)
    y = rng.normal(size=(12, 8)) * 2 + 2
    y[[3, 4, 8], :] = np.nan
    N, seed = 96, 11
    m, th = up.get_lorenz(1.0, 8, "rk4", 1), [8.0, 0.4]
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, 8, 8, N, 12, theta=[8.0, 0.4, 1.0], integrator=0, substeps=1)
    refs = [oracle.cpf(ocfg, y, seed=5, filter_id=50 + i)["trajectory"] for i in range(4)]
    g1 = up.ccpf_batch(N, m, th, y, refs[:2], seed=seed, ancestors=True)
    g2 = up.ccpf_batch(N, m, th, y, refs[:2], refs[2:], seed=seed, ancestors=True)
    ident = np.arange(N)
    for b in range(2):
        o1 = oracle.ccpf(ocfg, y, refs[b], seed=seed, filter_id=b, want_ancestors=True)
        o2 = oracle.ccpf(ocfg, y, refs[b], refs[2 + b], seed=seed, filter_id=b, want_ancestors=True)
        assert np.array_equal(g1["ancestors"][b, 0], o1["ancestors1"])
        assert np.array_equal(g2["ancestors"][b, 0], o2["ancestors1"]) and np.array_equal(g2["ancestors"][b, 1], o2["ancestors2"])
        for t in (0, 4, 5, 9):                                   # time 1 and the steps after an NA row: identity ancestors
            assert np.array_equal(g2["ancestors"][b, 0][t], ident) and np.array_equal(g2["ancestors"][b, 1][t], ident)
        assert np.allclose(g1["trajectory1"][b], o1["trajectory1"], rtol=1e-9, atol=1e-9)
        assert np.allclose(g2["trajectory1"][b], o2["trajectory1"], rtol=1e-9, atol=1e-9)
        assert np.allclose(g2["trajectory2"][b], o2["trajectory2"], rtol=1e-9, atol=1e-9)
        assert abs(g1["logz"][b] - o1["logz"]) <= 1e-9 * abs(o1["logz"])


def test_sharded_ccpf_estimator_single_process(up, golden):
    y = golden["lgssm_y"]
    est = up.unbiased_ccpf_estimator_sharded(up.get_ar(1), TH_AR, y, 64, 300, K=4, k=1, with_as=True, seed=3)
    g = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, 64, 300, K=4, k=1, with_as=True, seed=3)
    assert est["nreplicates"] == 300 and est["nfilters"] == g["nfilters"]
    assert np.allclose(est["mean"], g["hbar"].mean(axis=0), rtol=1e-12, atol=1e-12)
    assert np.allclose(est["se"], g["hbar"].std(axis=0, ddof=1) / math.sqrt(300), rtol=1e-9, atol=1e-12)


def test_rheeglynn_estimator_with_geometric_truncation(up, oracle, golden):
    """rheeglynn_estimator with lambda > 0 (R/rheeglynn_estimator.R:24-31, 58-59): the sum of differences stops at the geometric
    truncation variable and difference n is divided by the survival probability (1 - lambda)^n.  Expected values are rebuilt from
    the oracle's recorded coupled chains (samples1[n] = X_n, samples2[n-1] = X~_{n-1}) run to the same truncation."""
    import ctypes as C
    from unbiased_pimh_b200 import _lib as L
    y, th = golden["lgssm_y"][:40], [0.5, math.sqrt(10.0)]
    N, T, R, seed, lam = 32, 40, 12, 21, 0.2
    trunc = np.array([0, 1, 2, 3, 5, 8, 50, 50, 4, 1, 2, 50], dtype=np.int32)
    model = up.get_ar(1)
    cfg = model.config(N, th, nobs=T)
    hb, it, mt = np.empty((R, T + 1, 1)), np.empty(R, dtype=np.int32), np.empty(R, dtype=np.int32)
    o = L.ChainOut(); o.hbar, o.iteration, o.meetingtime = hb.ctypes.data, it.ctypes.data, mt.ctypes.data
    L.check(L.lib().cpimh_rheeglynn(C.byref(cfg), L.obs_matrix(y, T).ctypes.data_as(C.c_void_p), C.c_int64(R), C.c_uint64(seed), C.c_uint32(3),
                                    C.c_double(lam), L.ptr(trunc), C.c_int(-1), C.c_int(0), C.byref(o)))
    ocfg = oracle.make_config(oracle.MODEL_AR, 1, 1, N, T, theta=th)
    for r in range(R):
        if trunc[r] == 0:
            c = oracle.coupled_ccpf_chains(ocfg, y, seed, 3 + r, K=0, max_iterations=1)
            exp = c["samples1"][0]
        else:
            c = oracle.coupled_ccpf_chains(ocfg, y, seed, 3 + r, K=0, max_iterations=int(trunc[r]))
            exp = c["samples1"][0].copy()
            for n in range(1, c["iteration"] + 1):
                exp = exp + (c["samples1"][n] - c["samples2"][n - 1]) / (1.0 - lam) ** n
            assert it[r] == c["iteration"] <= trunc[r]
        assert np.allclose(hb[r, :, 0], exp, rtol=1e-9, atol=1e-9), r
    # the R-shaped front end draws the truncation itself and returns it
    e = up.rheeglynn_estimator(y, model, th, {"nparticles": N, "with_as": False, "lambda": 0.3}, seed=5)
    assert e["estimate"].shape == (1, T + 1) and e["truncation"] >= 0 and e["iteration"] <= max(e["truncation"], 0) and np.all(np.isfinite(e["estimate"]))


@pytest.mark.parametrize("name,N,with_as", [("ar1", 64, True), ("gss", 500, True), ("sv", 1000, False), ("ar3", 100, False)])
def test_conditional_filters_reference_order_mode_equals_the_fast_pass_and_the_oracle(up, oracle, golden, monkeypatch, name, N, with_as):
    """Exact ancestors in the conditional filters: the fast pass compares every chosen knot (of the CDFs of w, of mu = nu / alpha,
    of the two residuals, of the ancestor-sampling weights) and every coupling uniform against alpha with a rounding margin; a
    step that trips the check is repeated inside the kernel with alpha, mu, the residuals and all cumsums in the reference's order
    (src/coupledresamplingindexmatching.cpp:13-24, src/multinomial.cpp:17,40-41).  On random draws nothing is flagged; the
    forced reference-order mode must give the fast pass's ancestors and paths, and the oracle's."""
    if name == "ar1":
        m, th, y, oc = up.get_ar(1), TH_AR, golden["lgssm_y"][:50], (oracle.MODEL_AR, 1, 1)
    elif name == "ar3":
        m, th, y, oc = up.get_ar(3), TH_AR, np.random.default_rng(4).normal(size=(30, 3)), (oracle.MODEL_AR, 3, 3)
    elif name == "gss":
        m, th, y, oc = up.get_gss(), [], golden["gss_y"][:40], (oracle.MODEL_GSS, 1, 1)
    else:
        m, th, y, oc = up.get_levy_sv(), golden["sv_theta_mle"], golden["sv_y"][:30], (oracle.MODEL_LEVY_SV, 2, 1)
    cfg = oracle.make_config(oc[0], oc[1], oc[2], N, y.shape[0], theta=th)
    refs = _refs(oracle, cfg, y, 9, 4)
    fast1 = up.ccpf_batch(N, m, th, y, refs[:2], with_as=with_as, seed=31, first_filter=2, ancestors=True)
    fast2 = up.ccpf_batch(N, m, th, y, refs[:2], refs[2:], with_as=with_as, seed=32, first_filter=5, ancestors=True)
    assert up.lib().cpimh_last_exact_repeats() == 0
    monkeypatch.setenv("CPIMH_EXACT_MARGIN_SCALE", "1e9")           # many steps now trip the check and are repeated in reference order
    rep1 = up.ccpf_batch(N, m, th, y, refs[:2], with_as=with_as, seed=31, first_filter=2, ancestors=True)
    n1 = up.lib().cpimh_last_exact_repeats()
    rep2 = up.ccpf_batch(N, m, th, y, refs[:2], refs[2:], with_as=with_as, seed=32, first_filter=5, ancestors=True)
    assert n1 > 0 and up.lib().cpimh_last_exact_repeats() > 0
    monkeypatch.delenv("CPIMH_EXACT_MARGIN_SCALE")
    monkeypatch.setenv("CPIMH_CCPF_EXACT", "1")
    ex1 = up.ccpf_batch(N, m, th, y, refs[:2], with_as=with_as, seed=31, first_filter=2, ancestors=True)
    ex2 = up.ccpf_batch(N, m, th, y, refs[:2], refs[2:], with_as=with_as, seed=32, first_filter=5, ancestors=True)
    for f, e in ((fast1, ex1), (fast2, ex2), (rep1, ex1), (rep2, ex2)):
        for k in ("ancestors", "path_index", "trajectory1", "logz"):
            assert np.array_equal(f[k], e[k]), k
    assert np.array_equal(fast2["trajectory2"], ex2["trajectory2"]) and np.array_equal(rep2["trajectory2"], ex2["trajectory2"])
    for b in range(2):
        o = oracle.ccpf(cfg, y, refs[b], refs[2 + b], with_as=with_as, seed=32, filter_id=5 + b, want_ancestors=True)
        assert np.array_equal(ex2["ancestors"][b, 0], o["ancestors1"]) and np.array_equal(ex2["ancestors"][b, 1], o["ancestors2"])
        assert list(ex2["path_index"][b]) == o["path_index"]


def test_ccpf_chains_reference_order_mode(up, oracle, golden, monkeypatch):
    y = golden["lgssm_y"]
    kw = dict(K=5, k=2, with_as=True, seed=12, first_replicate=3, max_iterations=400)
    fast = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, 64, 60, **kw)
    monkeypatch.setenv("CPIMH_EXACT_MARGIN_SCALE", "1e9")
    rep = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, 64, 60, **kw)
    assert up.lib().cpimh_last_exact_repeats() > 0
    assert np.array_equal(fast["meetingtime"], rep["meetingtime"]) and np.array_equal(fast["hbar"], rep["hbar"])
    monkeypatch.delenv("CPIMH_EXACT_MARGIN_SCALE")
    monkeypatch.setenv("CPIMH_CCPF_EXACT", "1")
    ex = up.coupled_ccpf_chains_batch(up.get_ar(1), TH_AR, y, 64, 60, **kw)
    assert np.array_equal(fast["meetingtime"], ex["meetingtime"]) and np.array_equal(fast["iteration"], ex["iteration"])
    assert np.array_equal(fast["hbar"], ex["hbar"]) and fast["nfilters"] == ex["nfilters"]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=TH_AR)
    o = oracle.coupled_ccpf_batch(cfg, y, 12, 3, 60, k=2, K=5, with_as=True, max_iterations=400)
    assert np.array_equal(ex["meetingtime"], o["meetingtime"]) and np.allclose(ex["hbar"][:, 0, :], o["hbar"], rtol=1e-9, atol=1e-9)
