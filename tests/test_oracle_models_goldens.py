"""CPU tests: oracle model functions against numpy restatements, the Lorenz integrators against scipy, and the whole
oracle filter against the known answers / statistical goldens the reference's authors left behind."""
import math

import numpy as np
import pytest
from scipy.special import ndtr
from scipy.stats import norm

A_OBS = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
SK_THETA = [1.0] + list(A_OBS.ravel(order="F"))


def test_gss_model_functions(oracle):
    rng = np.random.default_rng(0)
    N = 50
    cfg = oracle.make_config(oracle.MODEL_GSS, 1, 1, N, 10)
    z = rng.normal(size=(1, N))
    x0 = oracle.model_rinit(cfg, z)
    assert np.array_equal(x0, math.sqrt(2.0) * z)
    for t in (1, 2, 7):
        z = rng.normal(size=(1, N))
        x1 = oracle.model_rtransition(cfg, x0, t, z)
        ref = 0.5 * x0 + 25 * x0 / (1 + x0 ** 2) + 8 * math.cos(1.2 * (t - 1)) + z
        assert np.allclose(x1, ref, rtol=1e-15, atol=1e-15)
    lw = oracle.model_dmeasurement(cfg, x0, [0.7])
    assert np.allclose(lw, norm.logpdf(0.7, loc=x0[0] ** 2 / 20, scale=math.sqrt(10)), rtol=1e-13)


def test_ar_model_functions(oracle):
    rng = np.random.default_rng(1)
    d, N = 5, 40
    cfg = oracle.make_config(oracle.MODEL_AR, d, d, N, 10, theta=[0.4, 1.7])
    A = oracle.create_A(0.4, d)
    z = rng.normal(size=(d, N))
    x0 = oracle.model_rinit(cfg, z)
    assert np.array_equal(x0, z)
    z = rng.normal(size=(d, N))
    x1 = oracle.model_rtransition(cfg, x0, 3, z)
    assert np.allclose(x1, A @ x0 + z, rtol=1e-14, atol=1e-14)
    y = rng.normal(size=d)
    lw = oracle.model_dmeasurement(cfg, x1, y)
    ref = -0.5 * np.sum(((x1 - y[:, None]) / 1.7) ** 2, axis=0) - d * math.log(1.7) - d * 0.9189385
    assert np.allclose(lw, ref, rtol=1e-13)


def test_levy_sv_model_functions(oracle):
    rng = np.random.default_rng(2)
    N = 30
    th = [0.2, -0.3, 0.8, 0.09, 0.05]
    cfg = oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, N, 10, theta=th)
    z0 = rng.gamma(3.0, 0.2, size=N); swe = rng.random(N) * 0.1; se = swe + rng.random(N) * 0.01
    x0 = oracle.model_rinit(cfg, np.stack([z0, swe, se]))
    z1 = math.exp(-0.05) * z0 + swe
    assert np.allclose(x0[1], z1, rtol=1e-15) and np.allclose(x0[0], (1 / 0.05) * (z0 - z1 + se), rtol=1e-12)
    swe2 = rng.random(N) * 0.1; se2 = swe2 + rng.random(N) * 0.01
    x1 = oracle.model_rtransition(cfg, x0, 4, np.stack([swe2, se2]))
    z2 = math.exp(-0.05) * x0[1] + swe2
    assert np.allclose(x1[1], z2, rtol=1e-15) and np.allclose(x1[0], (1 / 0.05) * (x0[1] - z2 + se2), rtol=1e-11)
    v = np.abs(x1[0]) + 0.01
    lw = oracle.model_dmeasurement(cfg, np.stack([v, x1[1]]), [0.3])
    assert np.allclose(lw, norm.logpdf(0.3, loc=0.2 - 0.3 * v, scale=np.sqrt(v)), rtol=1e-13)
    # Philox-driven noise: Poisson(lambda xi^2/w2) jump counts => mean of sum_e is rate * (w2/xi)
    cfgb = oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, 20000, 10, theta=th)
    nz = oracle.fill_trans_noise(cfgb, 17, 0, 3)
    rate = 0.05 * 0.8 ** 2 / 0.09
    assert abs(nz[1].mean() - rate * 0.09 / 0.8) < 4 * math.sqrt(rate * 2 * (0.09 / 0.8) ** 2 / 20000)
    assert abs((nz[1] == 0).mean() - math.exp(-rate)) < 0.012
    init = oracle.fill_init_noise(cfgb, 17, 0)
    shape, scale = 0.8 ** 2 / 0.09, 0.09 / 0.8
    assert abs(init[0].mean() - shape * scale) < 5 * math.sqrt(shape * scale ** 2 / 20000)
    assert abs(init[0].var() - shape * scale ** 2) < 0.1 * shape * scale ** 2


def ref_gillespie(x, E, U):
    """R/model_stochastic_kinetic.R:96-129 for one particle with its injected (E_m, U_m) sequence."""
    c = [0.1, 0.7, 0.35, 0.2, 0.1, 0.9, 0.3, 0.1]
    S = np.array([[0, 0, 1, 0, 0, 0, -1, 0], [0, 0, 0, 1, -2, 2, 0, -1], [-1, 1, 0, 0, 1, -1, 0, 0], [-1, 1, 0, 0, 0, 0, 0, 0]], dtype=float)
    x = x.copy(); t = 0.0; m = 0
    while True:
        h = np.array([c[0] * x[3] * x[2], c[1] * (5 - x[3]), c[2] * x[3], c[3] * x[0], c[4] * x[1] * (x[1] - 1) / 2, c[5] * x[2], c[6] * x[0], c[7] * x[1]])
        h0 = float(np.sum(h.astype(np.longdouble)))
        t = t + (1.0 / h0) * E[m]
        if not (t < 0.1):
            return x, m
        ev = int(np.sum(np.cumsum(h) < U[m] * h0))
        x = x + S[:, min(ev, 7)]
        m += 1


def test_stochastic_kinetic_model_functions(oracle):
    rng = np.random.default_rng(3)
    N, M = 64, 48
    cfg = oracle.make_config(oracle.MODEL_SK, 4, 2, N, 10, theta=SK_THETA, sk_max_events=M)
    x0 = oracle.model_rinit(cfg, np.zeros((1, N)))
    assert np.array_equal(x0, np.tile(np.array([[8.0], [8.0], [8.0], [5.0]]), (1, N)))
    x = x0
    total = 0
    for t in (1, 2, 3):
        E = rng.exponential(size=(M, N)); U = rng.random((M, N))
        x1 = oracle.model_rtransition(cfg, x, t, np.concatenate([E, U]))
        for n in range(N):
            r, m = ref_gillespie(x[:, n], E[:, n], U[:, n])
            assert np.array_equal(x1[:, n], r)
            total += m
        x = x1
    assert total > N                                  # events did happen
    assert np.all(x == np.round(x)) and np.all(x[3] >= 0) and np.all(x[3] <= 5)
    y = np.array([7.5, 21.0])
    lw = oracle.model_dmeasurement(cfg, x, y)
    ref = norm.logpdf(A_OBS @ x, loc=y[:, None], scale=1.0).sum(axis=0)
    assert np.allclose(lw, ref, rtol=1e-13)


def test_lorenz_model_functions(oracle):
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(4)
    for d in (8, 64):
        F = 8.0
        x0 = rng.normal(size=d) + 2

        def rhs(t, x):
            return np.roll(x, 1) * (np.roll(x, -1) - np.roll(x, 2)) - x + F
        ref = solve_ivp(rhs, (0, 0.05), x0, rtol=1e-12, atol=1e-12, method="DOP853").y[:, -1]
        x4 = oracle.lorenz_step(x0, F, oracle.INTEGRATOR_RK4, 4)
        x1 = oracle.lorenz_step(x0, F, oracle.INTEGRATOR_RK4, 1)
        assert np.max(np.abs(x4 - ref)) < 1e-7 and np.max(np.abs(x1 - ref)) < 2e-5
        assert np.max(np.abs(x4 - ref)) < np.max(np.abs(x1 - ref))
        if d == 8:
            xd = oracle.lorenz_step(x0, F, oracle.INTEGRATOR_DOPRI5)
            assert np.max(np.abs(xd - ref)) < 1e-5          # controller tolerance 1e-6 (abs + rel)
    N = 20
    cfg = oracle.make_config(oracle.MODEL_LORENZ, 8, 8, N, 5, theta=[8.0, 0.3, 0.7], integrator=oracle.INTEGRATOR_RK4, substeps=2)
    z = rng.normal(size=(8, N))
    x0 = oracle.model_rinit(cfg, z)
    assert np.allclose(x0, -1 + 4 * ndtr(z), rtol=1e-14)
    z = rng.normal(size=(8, N))
    x1 = oracle.model_rtransition(cfg, x0, 2, z)
    for n in range(N):
        assert np.allclose(x1[:, n], oracle.lorenz_step(x0[:, n], 8.0, 0, 2, 0.05, 0.10) + 0.3 * z[:, n], rtol=1e-14)
    assert np.all(oracle.model_dmeasurement(cfg, x1, np.full(8, np.nan)) == 0)     # NA observation (R/model_get_lorenz.R:39-40)


# ---------------------------------------------------------------- whole-filter goldens
def test_injected_noise_filter_matches_python_restatement(oracle, golden):
    """CPF(ref=NULL) (R/CPF.R:14-78) for get_gss with every random number injected, against a direct numpy restatement."""
    rng = np.random.default_rng(9)
    N, T = 64, 25
    y = golden["gss_y"][:T]
    cfg = oracle.make_config(oracle.MODEL_GSS, 1, 1, N, T, shuffle=1)
    noise = dict(init=rng.normal(size=(1, N)), trans=rng.normal(size=(T, 1, N)), resample=np.sort(rng.random((T, N)), axis=1),
                 shuffle=rng.random((T, N - 1)), path_u=rng.random(1))
    o = oracle.cpf(cfg, y, noise=noise, all_particles=True, all_paths=True, want_ancestors=True)
    x = math.sqrt(2.0) * noise["init"][0]
    w = np.full(N, 1.0 / N)
    hist_x, hist_a, lz = [x.copy()], [], []
    for t in range(1, T + 1):
        if t == 1:
            a = np.arange(N)
        else:
            cdf = np.cumsum(w); u = cdf[-1] * noise["resample"][t - 1]
            a = np.minimum(np.searchsorted(cdf, u, side="right"), N - 1)
            for i in range(1, N):
                j = min(int(math.floor(noise["shuffle"][t - 1, i - 1] * (i + 1))), i)
                a[i], a[j] = a[j], a[i]
        x = x[a]
        x = 0.5 * x + 25 * x / (1 + x * x) + 8 * math.cos(1.2 * (t - 1)) + noise["trans"][t - 1, 0]
        lw = norm.logpdf(y[t - 1, 0], loc=x * x / 20, scale=math.sqrt(10.0))
        m = lw.max(); ww = np.exp(lw - m); w = ww / ww.sum()
        lz.append(math.log(ww.mean()) + m)
        hist_x.append(x.copy()); hist_a.append(a.copy())
        assert np.array_equal(o["ancestors"][t - 1], a)
    assert np.allclose(o["logz_t"], lz, rtol=1e-12)
    k = int(np.searchsorted(np.cumsum(w), noise["path_u"][0], side="left"))
    assert o["path_index"] == min(k, N - 1)
    paths = np.empty((N, T + 1))
    for n in range(N):
        j = n
        for s in range(T, -1, -1):
            paths[n, s] = hist_x[s][j]
            if s > 0:
                j = hist_a[s - 1][j]
    assert np.allclose(o["trajectories"][0].T, paths, rtol=1e-12)
    assert np.allclose(o["trajectory"][0], paths[o["path_index"]], rtol=1e-12)
    assert np.allclose(o["exp_path"][0], (paths * w[:, None]).sum(axis=0), rtol=1e-11)


def test_kalman_known_answer_and_variance_table(oracle, golden):
    """log mean exp(log Z) -> exact Kalman log-likelihood of the shipped LGSSM data (with the reference's truncated
    density constant); Var(log Z) vs the authors' table lgssmlargem.RData (10^4 filters per N)."""
    y = golden["lgssm_y"]
    target = float(golden["lgssm_kalman_loglik_truncconst"])
    tab = dict(zip(golden["lgssm_varlogz_N"].astype(int).tolist(), golden["lgssm_varlogz"].tolist()))
    nrep = 6000
    for N in (10, 36, 150):
        cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, N, 100, theta=[0.5, math.sqrt(10.0)])
        lz, _ = oracle.cpf_batch(cfg, y, 2024, 0, nrep, want_traj=False)
        var = lz.var(ddof=1)
        # log Z is skewed: allow 5 "Gaussian" standard errors of both estimates
        se = tab[N] * math.sqrt(2.0 / (nrep - 1)) + tab[N] * math.sqrt(2.0 / 9999)
        assert abs(var - tab[N]) < 5 * se, (N, var, tab[N])
        if N == 150:
            m = lz.max(); lme = m + math.log(np.mean(np.exp(lz - m)))
            se_lme = math.sqrt(np.var(np.exp(lz - m)) / nrep) / np.mean(np.exp(lz - m))
            assert abs(lme - target) < 5 * se_lme + 1e-3, (lme, target)


def test_stochastic_kinetic_variance_golden(oracle, golden):
    """Var(log Z) at N=1000 on the shipped SK data vs stochastic_kinetic_mt_varz.RData (1.0222 with 10^4 filters)."""
    y = golden["sk_y"]
    cfg = oracle.make_config(oracle.MODEL_SK, 4, 2, 1000, 100, theta=SK_THETA)
    nrep = 400
    lz, _ = oracle.cpf_batch(cfg, y, 5, 0, nrep, want_traj=False)
    target = float(golden["sk_varlogz"][0])
    assert abs(lz.var(ddof=1) - target) < 5 * target * math.sqrt(2.0 / (nrep - 1)), lz.var(ddof=1)


def test_meeting_times_match_closed_form(oracle, golden):
    """P[tau = n] = int p(z)(1-p(z))^(n-1) phi(z; 0, s2) dz with p(z) = Phi(-z/s) + exp(-z + s2/2) Phi((z - s2)/s)
    (inst/reproduce/lgssm/lgssm_get_and_plot_meetingtimes.R:210-231) with s2 = Var(log Z) from the authors' table."""
    y = golden["lgssm_y"]
    N = 36
    s2 = float(golden["lgssm_varlogz"][list(golden["lgssm_varlogz_N"].astype(int)).index(N)]); s = math.sqrt(s2)
    z = np.linspace(-12 * s, 12 * s, 40001)
    p = norm.cdf(-z / s) + np.exp(-z + s2 / 2) * norm.cdf((z - s2) / s)
    p1 = float(np.sum(p * norm.pdf(z, scale=s)) * (z[1] - z[0]))
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, N, 100, theta=[0.5, math.sqrt(10.0)])
    nrep = 1500
    r = oracle.coupled_pimh_batch(cfg, y, 77, 0, nrep, k=0, K=1)
    f1 = float(np.mean(r["meetingtime"] == 1))
    assert abs(f1 - p1) < 4.5 * math.sqrt(p1 * (1 - p1) / nrep) + 0.02, (f1, p1)
    assert r["nfilters"] == int(np.sum(1 + r["iteration"]))
    assert np.all(r["iteration"] == np.maximum(r["meetingtime"], 1))


def test_hbar_matches_formula_and_smoothing_mean(oracle, golden):
    y = golden["lgssm_y"]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=[0.5, math.sqrt(10.0)])
    c = oracle.coupled_pimh_chains(cfg, y, 3, 5, K=4)
    assert c["finished"] and c["iteration"] >= 4 and c["samples1"].shape == (c["iteration"] + 1, 101)
    tau = c["meetingtime"]
    # once met, the chains coincide with a lag of one (X_t == Y_{t-1})
    for t in range(int(tau), c["iteration"] + 1):
        assert np.array_equal(c["samples1"][t], c["samples2"][t - 1])
    # unbiased estimator averages to the Kalman smoothing mean
    r = oracle.coupled_pimh_batch(cfg, y, 11, 0, 600, k=1, K=3)
    est = r["hbar"].mean(axis=0); se = r["hbar"].std(axis=0, ddof=1) / math.sqrt(600)
    # exact smoothing means by Kalman + RTS for x0~N(0,1), x_t = 0.5 x_{t-1} + N(0,1), y_t ~ N(x_t, 10)
    T = 100; a = 0.5; R = 10.0
    mp, Pp, mf, Pf = np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1)
    mf[0], Pf[0] = 0.0, 1.0
    for t in range(1, T + 1):
        mp[t], Pp[t] = a * mf[t - 1], a * a * Pf[t - 1] + 1
        K = Pp[t] / (Pp[t] + R)
        mf[t], Pf[t] = mp[t] + K * (y[t - 1, 0] - mp[t]), (1 - K) * Pp[t]
    ms = mf.copy()
    for t in range(T - 1, -1, -1):
        G = Pf[t] * a / Pp[t + 1]
        ms[t] = mf[t] + G * (ms[t + 1] - mp[t + 1])
    zscores = (est - ms) / se
    assert np.max(np.abs(zscores)) < 4.5 and np.mean(np.abs(zscores) > 2) < 0.15


def test_all_particles_chains_rao_blackwellised_estimator(oracle, golden):
    """coupled_pimh_chains with the *_all_particles kernels (R/pf_kernels.R:147-215): exp_samples follow the same accept
    decisions as samples, lag-one coincidence after meeting, and H_bar over them is unbiased with a smaller variance."""
    y = golden["lgssm_y"]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=[0.5, math.sqrt(10.0)])
    plain = oracle.coupled_pimh_chains(cfg, y, 3, 5, K=4, hbar_k=1)
    c = oracle.coupled_pimh_chains(cfg, y, 3, 5, K=4, all_particles=True, hbar_k=1)
    assert c["iteration"] == plain["iteration"] and c["meetingtime"] == plain["meetingtime"]
    assert np.array_equal(c["samples1"], plain["samples1"]) and np.array_equal(c["hbar"], plain["hbar"])
    assert c["exp_samples1"].shape == c["samples1"].shape and c["exp_samples2"].shape == c["samples2"].shape
    for t in range(int(c["meetingtime"]), c["iteration"] + 1):
        assert np.array_equal(c["exp_samples1"][t], c["exp_samples2"][t - 1])
    # rows of exp_samples1 change exactly when rows of samples1 change (accepted together)
    ch_s = np.any(np.diff(c["samples1"], axis=0) != 0, axis=1)
    ch_e = np.any(np.diff(c["exp_samples1"], axis=0) != 0, axis=1)
    assert np.array_equal(ch_s, ch_e)
    # first row = exp_path of the init filter = weighted average of its N paths
    f = oracle.cpf(cfg, y, seed=3, filter_id=5, all_particles=True, all_paths=True)
    assert np.allclose(c["exp_samples1"][0], f["exp_path"][0], rtol=0, atol=0)
    # variance reduction: over replicates the RB estimator has smaller spread at most time points, same mean within error
    hb, hbe = [], []
    for rep in range(150):
        r = oracle.coupled_pimh_chains(cfg, y, 13, rep, K=3, all_particles=True, hbar_k=1)
        hb.append(r["hbar"]); hbe.append(r["hbar_exp"])
    hb, hbe = np.array(hb), np.array(hbe)
    assert np.mean(hbe.var(axis=0) < hb.var(axis=0)) > 0.8
    z = (hb.mean(0) - hbe.mean(0)) / np.sqrt((hb.var(0) + hbe.var(0)) / 150)
    assert np.max(np.abs(z)) < 4.5


def test_conditional_filters_oracle_properties(oracle, golden):
    """oracle/ccpf.c: the reference particle is pinned, identical references give identical systems, dtransition equals
    scipy's normal log-density, and the coupled-CCPF estimator is unbiased for the Kalman smoothing mean."""
    from scipy.stats import norm
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=th)
    ref = oracle.cpf(cfg, y, seed=1, filter_id=0)["trajectory"]
    ref2 = oracle.cpf(cfg, y, seed=1, filter_id=1)["trajectory"]
    r = oracle.ccpf(cfg, y, ref, seed=2, filter_id=5, want_ancestors=True)
    assert np.all(r["ancestors1"][:, -1] == 63) and r["clamp_count"] == 0
    assert np.all(np.diff(r["ancestors1"][1:, :-1], axis=1) >= 0)          # N-1 sorted multinomial draws
    # if the pinned particle is selected at the end, the reference path is returned unchanged
    hits = 0
    for f in range(40):
        rr = oracle.ccpf(cfg, y, ref, seed=4, filter_id=f)
        if rr["path_index"][0] == 63:
            hits += 1
            assert rr["trajectory1"][0, -1] == ref[0, -1]
    same = oracle.ccpf(cfg, y, ref, ref, seed=2, filter_id=5)
    assert np.array_equal(same["trajectory1"], same["trajectory2"])
    two = oracle.ccpf(cfg, y, ref, ref2, seed=2, filter_id=5, with_as=True, want_ancestors=True)
    assert two["trajectory1"].shape == (1, 101) and np.all(two["ancestors1"] >= 0) and np.all(two["ancestors2"] < 64)
    # dtransition: gss and ar(1)
    x = np.linspace(-3, 3, 64)
    lt = oracle.model_dtransition(cfg, [0.7], x, 3)
    assert np.allclose(lt, norm.logpdf(0.7, loc=0.5 * x, scale=1.0), atol=2e-7)     # the reference truncates log(sqrt(2 pi)) to 0.9189385
    gcfg = oracle.make_config(oracle.MODEL_GSS, 1, 1, 64, 100)
    lg = oracle.model_dtransition(gcfg, [1.3], x, 4)
    mean = 0.5 * x + 25 * x / (1 + x * x) + 8 * math.cos(1.2 * 3)
    assert np.allclose(lg, norm.logpdf(1.3, loc=mean, scale=1.0), rtol=1e-12, atol=1e-12)
    # unbiasedness (ancestor sampling on: short meeting times)
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
    b = oracle.coupled_ccpf_batch(cfg, y, 11, 0, 400, k=3, K=8, with_as=True)
    est = b["hbar"].mean(axis=0); se = b["hbar"].std(axis=0, ddof=1) / math.sqrt(400)
    z = (est - ms) / se
    assert np.max(np.abs(z)) < 4.5 and np.mean(np.abs(z) > 2) < 0.15
    c = oracle.coupled_ccpf_chains(cfg, y, 11, 7, K=4, with_as=True, hbar_k=1)
    assert c["finished"] and c["samples1"].shape == (c["iteration"] + 1, 101) and c["samples2"].shape == (c["iteration"], 101)
    for t in range(int(c["meetingtime"]), c["iteration"] + 1):
        assert np.array_equal(c["samples1"][t], c["samples2"][t - 1])


def test_adaptive_pf_oracle(oracle, golden):
    """oracle/apf.c (pf(), R/CPF.R:106-155): ess_threshold = 1 resamples every step, a low threshold rarely; the likelihood
    estimator is unbiased for the Kalman value for multinomial and systematic resampling; loglik_t is the running sum."""
    y = golden["lgssm_y"]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 256, 100, theta=[0.5, math.sqrt(10.0)])
    kal = float(golden["lgssm_kalman_loglik"])
    for ess, systematic in ((1.0, False), (0.5, False), (0.5, True)):
        rs = [oracle.pf_adaptive(cfg, y, seed=3, filter_id=f, ess_threshold=ess, systematic=systematic, want_ancestors=(f == 0)) for f in range(120)]
        ll = np.array([r["loglik"] for r in rs])
        lme = math.log(np.mean(np.exp(ll - ll.max()))) + ll.max()
        assert abs(lme - kal) < 0.12, (ess, systematic, lme)
        frac = np.mean([r["resampled"].mean() for r in rs])
        assert frac == 1.0 if ess == 1.0 else 0.02 < frac < 0.6
        r0 = rs[0]
        assert r0["loglik_t"][-1] == r0["loglik"] and np.all(np.diff(r0["loglik_t"]) < 0) and r0["clamp_count"] == 0
        a = r0["ancestors"]
        for t in range(100):
            if r0["resampled"][t]:
                assert np.all(np.diff(a[t]) >= 0) and a[t].min() >= 0 and a[t].max() < 256      # sorted draws / systematic: ascending
            else:
                assert np.array_equal(a[t], np.arange(256))
        assert r0["pf_means"].shape == (100, 1) and r0["path"].shape == (101, 1)
