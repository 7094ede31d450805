"""GPU tests of the coupled-PIMH driver and full-size property checks of the hot path (BASELINE configs)."""
import math

import numpy as np
import pytest
from scipy.stats import norm

pytestmark = pytest.mark.gpu

A_OBS = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
SK_THETA = {"A_obs": A_OBS, "s2_obs": 1.0}


@pytest.fixture(scope="module")
def up():
    import unbiased_pimh_b200 as m
    return m


def test_coupled_chains_match_oracle_replicate_for_replicate(up, oracle, golden):
    """Same Philox streams => same proposals, same accept/reject decisions, same meeting times, H_bar within 1e-10."""
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    for (k, K) in ((0, 1), (1, 4)):
        r = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 64, 96, K=K, k=k, seed=31, first_replicate=7, record_samples=16)
        cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=th)
        o = oracle.coupled_pimh_batch(cfg, y, 31, 7, 96, k=k, K=K)
        assert np.array_equal(r["meetingtime"], o["meetingtime"]) and np.array_equal(r["iteration"], o["iteration"])
        assert r["nfilters"] == o["nfilters"]
        assert np.allclose(r["hbar"][:, 0, :], o["hbar"], rtol=1e-10, atol=1e-10)
        assert np.allclose(r["sum_hbar"][0], o["hbar"].sum(axis=0), rtol=1e-10, atol=1e-9)
        for rep in (0, 5, 40):
            c = oracle.coupled_pimh_chains(cfg, y, 31, 7 + rep, K=K)
            g = r["c_chains"][rep]
            n = min(c["iteration"], 16)
            assert g["meetingtime"] == c["meetingtime"] and g["iteration"] == c["iteration"] and g["finished"]
            assert np.allclose(g["samples1"], c["samples1"][:n + 1], rtol=1e-10, atol=1e-10)
            assert np.allclose(g["samples2"], c["samples2"][:n], rtol=1e-10, atol=1e-10)
            assert np.allclose(g["log_pdf1"][:, 0], c["log_pdf1"][:n + 1], rtol=1e-10)
            if c["iteration"] <= 16:
                assert np.allclose(up.H_bar(g, k=k, K=K), r["hbar"][rep, 0], rtol=1e-10, atol=1e-10)
                assert np.allclose(up.BC(g, k=k, K=K), r["bc"][rep, 0], rtol=1e-10, atol=1e-10)


def test_multi_component_models_in_chains(up, oracle, golden):
    th = golden["sv_theta_mle"]
    y = golden["sv_y"][:50]
    r = up.coupled_pimh_chains_batch(up.get_levy_sv(), th, y, 512, 24, K=2, k=0, seed=8)
    cfg = oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, 512, 50, theta=th)
    o = oracle.coupled_pimh_batch(cfg, y, 8, 0, 24, k=0, K=2)
    assert np.array_equal(r["meetingtime"], o["meetingtime"])
    assert r["hbar"].shape == (24, 2, 51) and np.allclose(r["hbar"][:, 0, :], o["hbar"], rtol=1e-9, atol=1e-10)


def test_r_shaped_chain_driver_on_gpu(up, golden):
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    up.set_seed(3)
    ker = up.get_pimh_kernels(False, model=up.get_ar(1), theta=th, observations=y)
    c = up.coupled_pimh_chains(ker["single_kernel"], ker["coupled_kernel"], ker["pimh_init"], 100, K=3)
    assert c["finished"] and c["iteration"] >= 3 and c["samples1"].shape == (c["iteration"] + 1, 101)
    hb = up.H_bar(c, k=1, K=3)
    assert hb.shape == (101,) and np.all(np.isfinite(hb))
    kap = up.get_pimh_kernels(False, all_particles=True, model=up.get_ar(1), theta=th, observations=y)
    c2 = up.coupled_pimh_chains(kap["single_kernel"], kap["coupled_kernel"], kap["pimh_init"], 100, K=2)
    assert c2["exp_samples1"].shape == c2["samples1"].shape
    e = up.rheeglynn_estimator(y, up.get_ar(1), th, {"nparticles": 100, "lambda": 0, "with_as": False}, seed=4)
    assert e["estimate"].shape == (1, 101) and e["iteration"] >= 1


def test_statistical_goldens_lgssm(up, golden):
    """Authors' Var(log Z) table (lgssmlargem.RData), the Kalman log-likelihood and the closed-form meeting-time law."""
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    nrep = 10000
    tab = dict(zip(golden["lgssm_varlogz_N"].astype(int).tolist(), golden["lgssm_varlogz"].tolist()))
    target = float(golden["lgssm_kalman_loglik_truncconst"])
    for N in (10, 21, 59, 150):
        lz = up.cpf_batch(N, up.get_ar(1), th, y, nrep, seed=100 + N, store_paths=0)["logz"]
        se = 2 * tab[N] * math.sqrt(2.0 / (nrep - 1))
        assert abs(lz.var(ddof=1) - tab[N]) < 5 * se, (N, lz.var(ddof=1), tab[N])
        if N >= 59:
            m = lz.max(); w = np.exp(lz - m)
            assert abs(m + math.log(w.mean()) - target) < 5 * w.std() / w.mean() / math.sqrt(nrep) + 1e-3
    N = 36
    s2 = tab[N]; s = math.sqrt(s2)
    z = np.linspace(-12 * s, 12 * s, 40001)
    p = norm.cdf(-z / s) + np.exp(-z + s2 / 2) * norm.cdf((z - s2) / s)
    R = 20000
    mt = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, N, R, K=1, seed=55)["meetingtime"]
    for n in (1, 2, 3):
        pn = float(np.sum(p * (1 - p) ** (n - 1) * norm.pdf(z, scale=s)) * (z[1] - z[0]))
        assert abs(np.mean(mt == n) - pn) < 4.5 * math.sqrt(pn * (1 - pn) / R) + 0.01, (n, np.mean(mt == n), pn)


def test_statistical_goldens_stochastic_kinetic(up, golden):
    """stochastic_kinetic_mt_varz.RData: Var(log Z) and meeting-time frequencies at N = 1000, 3000 (10^4 replicates there)."""
    y = golden["sk_y"]
    for idx, N in ((0, 1000), (2, 3000)):
        nrep = 4000
        lz = up.cpf_batch(N, up.get_stochastic_kinetic(), SK_THETA, y, nrep, seed=7 + N, store_paths=0)["logz"]
        target = float(golden["sk_varlogz"][idx])
        assert abs(lz.var(ddof=1) - target) < 5 * target * (math.sqrt(2.0 / (nrep - 1)) + math.sqrt(2.0 / 9999)), (N, lz.var(ddof=1), target)
        R = 3000
        mt = up.coupled_pimh_chains_batch(up.get_stochastic_kinetic(), SK_THETA, y, N, R, K=1, seed=9 + N)["meetingtime"]
        hist = golden["sk_mt_hist"][idx]; p1 = hist[1] / hist.sum()
        assert abs(np.mean(mt == 1) - p1) < 4.5 * math.sqrt(p1 * (1 - p1) * (1 / R + 1 / hist.sum())), (N, np.mean(mt == 1), p1)
        assert abs(mt.mean() - float(golden["sk_mt_mean"][idx])) < 0.12


def test_smoothing_mean_within_monte_carlo_error(up, golden):
    """End-to-end: unbiased smoothing estimator vs the exact Kalman/RTS smoother of the LGSSM data."""
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    R = 4000
    r = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 128, R, K=5, k=2, seed=77)
    est = r["hbar"][:, 0, :].mean(axis=0); se = r["hbar"][:, 0, :].std(axis=0, ddof=1) / math.sqrt(R)
    res = up.combine_sums(np.concatenate([r["sum_hbar"].T.ravel(), r["sum_hbar_sq"].T.ravel(), [r["nfilters"], R]]), 1, 100)
    assert np.allclose(res["mean"][0], est, rtol=1e-10) and np.allclose(res["se"][0], se, rtol=1e-6)
    T = 100; a = 0.5; Rv = 10.0
    mp, Pp, mf, Pf = np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1), np.zeros(T + 1)
    Pf[0] = 1.0
    for t in range(1, T + 1):
        mp[t], Pp[t] = a * mf[t - 1], a * a * Pf[t - 1] + 1
        K = Pp[t] / (Pp[t] + Rv)
        mf[t], Pf[t] = mp[t] + K * (y[t - 1, 0] - mp[t]), (1 - K) * Pp[t]
    ms = mf.copy()
    for t in range(T - 1, -1, -1):
        ms[t] = mf[t] + Pf[t] * a / Pp[t + 1] * (ms[t + 1] - mp[t + 1])
    zs = (est - ms) / se
    assert np.max(np.abs(zs)) < 4.5 and np.mean(np.abs(zs) > 2) < 0.12


def test_sharding_is_independent_of_the_number_of_gpus(up, golden):
    """Replicate r always uses Philox stream r: one shard of 64 == two shards of 32 == the same hbar, bit for bit."""
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    full = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 50, 64, K=1, seed=5)
    parts = []
    for g in range(2):
        lo, hi = up.shard_range(64, g, 2)
        parts.append(up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 50, hi - lo, K=1, seed=5, first_replicate=lo))
    assert np.array_equal(full["hbar"], np.concatenate([p["hbar"] for p in parts]))
    assert np.array_equal(full["meetingtime"], np.concatenate([p["meetingtime"] for p in parts]))
    assert full["nfilters"] == sum(p["nfilters"] for p in parts)
    one = up.unbiased_estimator_sharded(up.get_ar(1), th, y, 50, 64, K=1, seed=5)       # world size 1: no collective
    assert np.allclose(one["mean"][0], full["hbar"][:, 0, :].mean(axis=0), rtol=1e-10) and one["nreplicates"] == 64


def test_full_size_config2_properties(up, golden):
    """BASELINE configs[1] at full size (levy_sv, T=1000, N=4096): determinism, stream independence of the batch layout,
    finite log Z, and the estimator's mean log Z stable across seeds."""
    th = golden["sv_theta_mle"]
    y = np.tile(golden["sv_y"], (2, 1))
    B = 64
    pb = up.PfBatch(up.get_levy_sv(), th, y, 4096, B)
    pb.run(B, 17)
    a = pb.fetch(path_index=True)
    pb.run(B, 17)
    b = pb.fetch(path_index=True)
    assert np.array_equal(a["logz"], b["logz"]) and np.array_equal(a["trajectory"], b["trajectory"])     # bit-reproducible
    pb.run(32, 17, first_filter=32)
    c = pb.fetch()
    assert np.array_equal(c["logz"], a["logz"][32:]) and np.array_equal(c["trajectory"], a["trajectory"][32:])
    assert np.all(np.isfinite(a["logz"])) and np.all(a["trajectory"][:, 0, :] > 0)                        # variance component stays positive
    assert np.all((a["path_index"] >= 0) & (a["path_index"] < 4096))
    pb.run(B, 18)
    d = pb.fetch()
    assert abs(a["logz"].mean() - d["logz"].mean()) < 5 * math.sqrt((a["logz"].var() + d["logz"].var()) / B)
    pb.close()


def test_full_size_config3_ancestors_sorted_and_in_range(up, golden):
    """BASELINE configs[2] shape (stochastic kinetic, N=8192): recorded ancestors are a non-decreasing sequence in range
    (multinomial from sorted uniforms, shuffle skipped), states stay on the integer lattice with 0 <= DNA <= 5."""
    y = golden["sk_y"]
    pb = up.PfBatch(up.get_stochastic_kinetic(), SK_THETA, y, 8192, 2)
    pb.record_ancestors(True)
    pb.run(2, 3)
    r = pb.fetch()
    anc = pb.fetch_ancestors(1)
    assert anc.min() >= 0 and anc.max() < 8192 and np.all(np.diff(anc, axis=1) >= 0) and np.array_equal(anc[0], np.arange(8192))
    tr = r["trajectory"]
    assert np.all(tr == np.round(tr)) and np.all(tr[:, 3, :] >= 0) and np.all(tr[:, 3, :] <= 5)
    assert np.array_equal(tr[:, :, 0], np.tile([8.0, 8.0, 8.0, 5.0], (2, 1)))
    pb.close()


def test_round_based_driver_equals_persistent_kernel(up, oracle, golden):
    """The two execution strategies of the chain driver (persistent CTA per replicate / rounds of batched, speculative
    proposals) consume the same Philox streams in the same order: identical meeting times, iterations, H_bar, filters."""
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    for (N, R, k, K, mi) in ((20, 3000, 0, 1, math.inf), (64, 500, 1, 4, math.inf), (10, 800, 0, 2, 6)):
        a = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, N, R, K=K, k=k, max_iterations=mi, seed=12, first_replicate=3, mode=1)
        b = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, N, R, K=K, k=k, max_iterations=mi, seed=12, first_replicate=3, mode=2)
        assert np.array_equal(a["meetingtime"], b["meetingtime"]) and np.array_equal(a["iteration"], b["iteration"])
        assert a["nfilters"] == b["nfilters"] == int(np.sum(1 + a["iteration"]))
        assert np.array_equal(a["hbar"], b["hbar"])
        assert np.allclose(a["sum_hbar"], b["sum_hbar"], rtol=1e-13)
        if mi == 6:
            assert a["iteration"].max() <= 6 and np.any(a["meetingtime"] < 0)      # hit max_iterations before meeting
            # such replicates are reported as not finished (c_chains$finished, R/debiasing_algorithm.R:262) and stay OUT of the
            # sums the estimator averages (H_bar refuses a chain with K > iteration); round 1 averaged their truncated H_bar in
            for r in (a, b):
                fin = r["finished"]
                assert np.array_equal(fin, r["meetingtime"] > 0) and r["nfinished"] == int(fin.sum()) and 0 < r["nfinished"] < R
                assert np.allclose(r["sum_hbar"], r["hbar"][fin].sum(axis=0), rtol=1e-12, atol=1e-12)
        else:
            assert a["finished"].all() and a["nfinished"] == R
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 20, 100, theta=th)
    o = oracle.coupled_pimh_batch(cfg, y, 12, 3, 200, k=0, K=1)
    assert np.array_equal(b["meetingtime"][:0], o["meetingtime"][:0])
    r = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 20, 200, K=1, k=0, seed=12, first_replicate=3, mode=2)
    assert np.array_equal(r["meetingtime"], o["meetingtime"]) and np.allclose(r["hbar"][:, 0, :], o["hbar"], rtol=1e-10, atol=1e-10)
    cb = up.ChainBatch(up.get_ar(1), th, y, 20, 3000)
    cb.run(3000, 1)
    out = cb.fetch(hbar=False)
    assert cb.filters_launched >= out["nfilters"]                                    # speculative proposals may be discarded
    cb.close()


def test_chains_with_tree_path_storage(up, golden):
    """Both chain strategies on top of the Jacob-Murray-Rubenthaler tree instead of the dense history."""
    y = golden["lgssm_y"]
    th = [0.5, math.sqrt(10.0)]
    ref = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 100, 300, K=2, seed=4)
    for mode in (1, 2):
        r = up.coupled_pimh_chains_batch(up.get_ar(1), th, y, 100, 300, K=2, seed=4, mode=mode, store_paths=3)
        assert np.array_equal(r["meetingtime"], ref["meetingtime"]) and np.array_equal(r["hbar"], ref["hbar"])


def test_all_particles_batch_and_chains_match_oracle(up, oracle, golden):
    """all_particles = TRUE end to end: the in-kernel weighted-average path of every filter of a batch equals the oracle's
    pf_get_weighted_average, for both path stores, and the chain driver's H_bar over exp_samples equals the oracle's."""
    from unbiased_pimh_b200.filters import PfBatch
    y = golden["sv_y"][:40]
    th = golden["sv_theta_mle"]
    ocfg = oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, 300, 40, theta=th)
    for sp in (2, 3):
        pf = PfBatch(up.get_levy_sv(), th, y, 300, 20, store_paths=sp)
        pf.all_particles(True)
        pf.run(20, 77, first_filter=3)
        ep = pf.fetch_exp_paths()
        one = pf.fetch_all_particles(b=11)
        pf.close()
        assert ep.shape == (20, 2, 41)
        assert np.allclose(ep[11], one["exp_path"], rtol=1e-12, atol=1e-13)
        for b in (0, 11, 19):
            o = oracle.cpf(ocfg, y, seed=77, filter_id=3 + b, all_particles=True)
            assert np.allclose(ep[b], o["exp_path"], rtol=1e-9, atol=1e-10)
    yl = golden["lgssm_y"]
    tha = [0.5, math.sqrt(10.0)]
    cfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 100, theta=tha)
    for (k, K) in ((0, 1), (1, 4)):
        r = up.coupled_pimh_chains_batch(up.get_ar(1), tha, yl, 64, 48, K=K, k=k, seed=31, first_replicate=7, all_particles=True)
        plain = up.coupled_pimh_chains_batch(up.get_ar(1), tha, yl, 64, 48, K=K, k=k, seed=31, first_replicate=7)
        assert np.array_equal(r["meetingtime"], plain["meetingtime"]) and np.allclose(r["hbar"], plain["hbar"], rtol=0, atol=0)
        for rep in range(48):
            c = oracle.coupled_pimh_chains(cfg, yl, 31, 7 + rep, K=K, all_particles=True, hbar_k=k)
            assert c["meetingtime"] == r["meetingtime"][rep]
            assert np.allclose(r["hbar_rb"][rep, 0], c["hbar_exp"], rtol=1e-10, atol=1e-10)
        assert np.allclose(r["sum_hbar_rb"], r["hbar_rb"].sum(axis=0), rtol=1e-12, atol=1e-10)
        assert np.allclose(r["sum_hbar_rb_sq"], (r["hbar_rb"] ** 2).sum(axis=0), rtol=1e-12, atol=1e-10)
    # a multi-component model through the tree store
    r = up.coupled_pimh_chains_batch(up.get_levy_sv(), th, y, 256, 12, K=2, seed=8, all_particles=True, store_paths=3)
    cfg = oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, 256, 40, theta=th)
    for rep in range(12):
        c = oracle.coupled_pimh_chains(cfg, y, 8, rep, K=2, all_particles=True, hbar_k=0)
        assert np.allclose(r["hbar_rb"][rep, 0], c["hbar_exp"], rtol=1e-9, atol=1e-10)
