"""GPU parity of the fused filter kernel (through the C ABI) against the CPU oracle.

Tolerances (fp64): per-step log Z within 1e-10 relative given identical injected noise (north_star); ancestor
indices identical; trajectories within 1e-9 absolute (AR/Lorenz amplify rounding along the path).
"""
import math

import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A_OBS = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
RTOL_LOGZ = 1e-10


@pytest.fixture(scope="module")
def up():
    import unbiased_pimh_b200 as m
    return m


def model_cases(up, oracle, golden):
    rng = np.random.default_rng(0)
    th_sv = golden["sv_theta_mle"]
    return {
        "gss": dict(model=up.get_gss(), theta=[], oid=oracle.MODEL_GSS, otheta=[], obs=golden["gss_y"], d=1, dy=1, okw={}),
        "ar1": dict(model=up.get_ar(1), theta=[0.5, math.sqrt(10)], oid=oracle.MODEL_AR, otheta=[0.5, math.sqrt(10)], obs=golden["lgssm_y"], d=1, dy=1, okw={}),
        "ar5": dict(model=up.get_ar(5), theta=[0.3, 1.5], oid=oracle.MODEL_AR, otheta=[0.3, 1.5], obs=rng.normal(size=(40, 5)) * 2, d=5, dy=5, okw={}),
        "ar10": dict(model=up.get_ar(10), theta=[0.3, math.sqrt(10)], oid=oracle.MODEL_AR, otheta=[0.3, math.sqrt(10)], obs=rng.normal(size=(30, 10)) * 2, d=10, dy=10, okw={}),
        "levy_sv": dict(model=up.get_levy_sv(), theta=th_sv, oid=oracle.MODEL_LEVY_SV, otheta=th_sv, obs=golden["sv_y"][:60], d=2, dy=1, okw={}),
        "sk": dict(model=up.get_stochastic_kinetic(), theta={"A_obs": A_OBS, "s2_obs": 1.0}, oid=oracle.MODEL_SK, otheta=[1.0] + list(A_OBS.ravel(order="F")),
                   obs=golden["sk_y"][:40], d=4, dy=2, okw={"sk_max_events": 48}),
        "lorenz_rk4": dict(model=up.get_lorenz(1.0, 8, "rk4", 2), theta=[8.0, 0.5], oid=oracle.MODEL_LORENZ, otheta=[8.0, 0.5, 1.0], obs=rng.normal(size=(20, 8)) * 3 + 2,
                           d=8, dy=8, okw={"integrator": 0, "substeps": 2}),
        "lorenz_dopri5": dict(model=up.get_lorenz(1.0, 8, "dopri5"), theta=[8.0, 0.5], oid=oracle.MODEL_LORENZ, otheta=[8.0, 0.5, 1.0], obs=rng.normal(size=(20, 8)) * 3 + 2,
                              d=8, dy=8, okw={"integrator": 1}),
    }


NAMES = ["gss", "ar1", "ar5", "ar10", "levy_sv", "sk", "lorenz_rk4", "lorenz_dopri5"]


def injected_noise(rng, oracle, ocfg, B, N, T, shuffle):
    ni, nt = oracle.n_init(ocfg), oracle.n_trans(ocfg)
    nz = dict(resample=np.sort(rng.random((B, T, N)), axis=2), path_u=rng.random(B))
    if ocfg.model == oracle.MODEL_LEVY_SV:
        z0 = rng.gamma(7.0, 0.11, size=(B, 1, N)); sw = rng.exponential(0.02, size=(B, 1, N)) * (rng.random((B, 1, N)) < 0.3)
        nz["init"] = np.concatenate([z0, sw, sw * 1.02], axis=1)
        swt = rng.exponential(0.02, size=(B, T, 1, N)) * (rng.random((B, T, 1, N)) < 0.3)
        nz["trans"] = np.concatenate([swt, swt * 1.03], axis=2)
    elif ocfg.model == oracle.MODEL_SK:
        M = nt // 2
        nz["trans"] = np.concatenate([rng.exponential(size=(B, T, M, N)), rng.random((B, T, M, N))], axis=2)
    else:
        nz["init"] = rng.normal(size=(B, ni, N))
        nz["trans"] = rng.normal(size=(B, T, nt, N))
    if shuffle:
        nz["shuffle"] = rng.random((B, T, N - 1))
    return nz


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("shuffle", [0, 1])
@pytest.mark.parametrize("store_paths", [2, 3])
def test_injected_noise_parity(up, oracle, golden, name, shuffle, store_paths):
    """Identical pre-generated noise into both implementations: ancestors identical, per-step log Z within 1e-10."""
    c = model_cases(up, oracle, golden)[name]
    rng = np.random.default_rng(zlib.crc32(("%s-%d" % (name, shuffle)).encode()))      # deterministic across processes (str hashes are salted)
    N, B = (96 if c["d"] >= 8 else 200), 3
    T = c["obs"].shape[0]
    ocfg = oracle.make_config(c["oid"], c["d"], c["dy"], N, T, theta=c["otheta"], shuffle=shuffle, **c["okw"])
    nz = injected_noise(rng, oracle, ocfg, B, N, T, shuffle)
    pb = up.PfBatch(c["model"], c["theta"], c["obs"], N, B, shuffle=shuffle, store_paths=store_paths, sk_max_events=c["okw"].get("sk_max_events", 64))
    pb.record_ancestors(True)
    pb.set_noise(B, nz)
    pb.run(B, 0)
    r = pb.fetch(logz_t=True, path_index=True)
    for b in range(B):
        o = oracle.cpf(ocfg, c["obs"], noise={k: v[b] for k, v in nz.items()}, all_particles=True, all_paths=(b == 0), want_ancestors=True)
        assert np.array_equal(pb.fetch_ancestors(b), o["ancestors"]), (name, b)
        assert np.max(np.abs(r["logz_t"][b] - o["logz_t"]) / np.abs(o["logz_t"])) < RTOL_LOGZ
        assert abs(r["logz"][b] - o["logz"]) < RTOL_LOGZ * abs(o["logz"])
        assert int(r["path_index"][b]) == o["path_index"]
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)
        ap = pb.fetch_all_particles(b, trajectories=(b == 0))
        assert np.allclose(ap["weights"], o["normweights"], rtol=1e-9, atol=1e-300) and abs(ap["weights"].sum() - 1) < 1e-12
        assert np.allclose(ap["exp_path"], o["exp_path"], rtol=1e-9, atol=1e-9)
        if b == 0:
            assert np.allclose(ap["trajectories"], o["trajectories"], rtol=1e-9, atol=1e-9)
    pb.close()


@pytest.mark.parametrize("name,N", [("gss", 256), ("gss", 1000), ("ar1", 150), ("ar10", 1024), ("levy_sv", 4096), ("sk", 2048), ("sk", 8192),
                                    ("lorenz_rk4", 512), ("lorenz_dopri5", 300)])
def test_philox_seed_for_seed_parity(up, oracle, golden, name, N):
    """No injected arrays: both sides draw from the same Philox streams (bit-identical uniforms), so whole filters agree
    up to libm-vs-CUDA rounding of log/exp/sincos."""
    c = model_cases(up, oracle, golden)[name]
    T = c["obs"].shape[0]
    B, seed = 3, 4242
    r = up.cpf_batch(N, c["model"], c["theta"], c["obs"], B, seed=seed, first_filter=5, logz_t=True)
    okw = {k: v for k, v in c["okw"].items() if k != "sk_max_events"}
    ocfg = oracle.make_config(c["oid"], c["d"], c["dy"], N, T, theta=c["otheta"], **okw)
    for b in range(B):
        o = oracle.cpf(ocfg, c["obs"], seed=seed, filter_id=5 + b)
        assert np.max(np.abs(r["logz_t"][b] - o["logz_t"]) / np.abs(o["logz_t"])) < RTOL_LOGZ, (name, b)
        assert int(r["path_index"][b]) == o["path_index"]
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-8, atol=1e-8)


def test_dense_and_tree_path_storage_agree(up, golden):
    th = golden["sv_theta_mle"]
    y = np.tile(golden["sv_y"], (2, 1))[:300]
    out = {}
    for sp in (2, 3):
        pb = up.PfBatch(up.get_levy_sv(), th, y, 1024, 8, store_paths=sp, tree_capacity=24 * 1024)
        assert pb.effective_config()["path_storage"] == {2: "dense", 3: "tree"}[sp]
        pb.run(8, 99)
        out[sp] = pb.fetch(logz_t=True, path_index=True)
        out[sp]["ap"] = pb.fetch_all_particles(3, trajectories=True)
        pb.close()
    for k in ("logz", "logz_t", "trajectory", "path_index"):
        assert np.array_equal(out[2][k], out[3][k]), k
    assert np.array_equal(out[2]["ap"]["trajectories"], out[3]["ap"]["trajectories"])
    assert np.allclose(out[2]["ap"]["exp_path"], out[3]["ap"]["exp_path"], rtol=1e-13)
    pb = up.PfBatch(up.get_levy_sv(), th, y, 1024, 8, store_paths=0)          # log Z only
    pb.run(8, 99)
    assert np.array_equal(pb.fetch(trajectory=False)["logz"], out[2]["logz"])
    pb.close()


def test_tree_overflow_is_reported_and_retried(up, golden):
    y = golden["lgssm_y"]
    m = up.get_ar(1)
    pb = up.PfBatch(m, [0.5, 1e6], y, 512, 4, store_paths=3, tree_capacity=3 * 512)   # flat weights (huge sigma_y) => large genealogy; minimum capacity 3N
    pb.run(4, 1)
    with pytest.raises(up.CpimhError) as ei:
        pb.fetch()
    assert ei.value.code == 3 and "capacity" in str(ei.value)
    pb.close()
    a = up.cpf_batch(512, m, [0.5, 1e6], y, 4, seed=1, store_paths=3, tree_capacity=3 * 512)   # one-shot API doubles the capacity and reruns
    b = up.cpf_batch(512, m, [0.5, 1e6], y, 4, seed=1, store_paths=2)
    assert np.array_equal(a["logz"], b["logz"]) and np.array_equal(a["trajectory"], b["trajectory"])


def test_na_observations_skip_resampling(up, oracle):
    """R/CPF.R:38-40: no resampling after a row whose first entry is NA; get_lorenz's dmeasurement returns 0 for NA rows."""
    rng = np.random.default_rng(3)
    y = rng.normal(size=(12, 8)) * 2 + 2
    y[[3, 4, 8], :] = np.nan
    N, seed = 200, 11
    r = up.cpf_batch(N, up.get_lorenz(1.0, 8, "rk4", 1), [8.0, 0.4], y, 2, seed=seed, logz_t=True)
    ocfg = oracle.make_config(oracle.MODEL_LORENZ, 8, 8, N, 12, theta=[8.0, 0.4, 1.0], integrator=0, substeps=1)
    for b in range(2):
        o = oracle.cpf(ocfg, y, seed=seed, filter_id=b)
        assert np.allclose(r["logz_t"][b], o["logz_t"], rtol=RTOL_LOGZ, atol=1e-300)
        assert np.all(r["logz_t"][b][[3, 4, 8]] == 0.0)
        assert int(r["path_index"][b]) == o["path_index"]


def test_r_shaped_single_filter_api(up, oracle, golden):
    y = golden["gss_y"]
    up.set_seed(17)
    r = up.CPF(256, up.get_gss(), [], y)
    assert r["trajectory"].shape == (1, 101) and np.isfinite(r["logz"])
    r2 = up.CPF(256, up.get_gss(), [], y, all_particles=True, seed=5)
    pf = up.particlefilter(256, up.get_gss(), [], y, seed=5)
    assert pf["trajectories"].shape == (1, 101, 256) and abs(pf["weights"].sum() - 1) < 1e-12
    assert np.allclose(r2["exp_path"], up.pf_get_weighted_average(pf["trajectories"], pf["weights"]), rtol=1e-12)
    ocfg = oracle.make_config(oracle.MODEL_GSS, 1, 1, 256, 100)
    o = oracle.cpf(ocfg, y, seed=5, filter_id=0, all_particles=True, all_paths=True)
    assert np.allclose(pf["trajectories"], o["trajectories"], rtol=1e-9, atol=1e-9) and np.allclose(r2["exp_path"], o["exp_path"], rtol=1e-9)


def test_model_closures_match_oracle(up, oracle, golden):
    """model$rinit / rtransition / dmeasurement one call at a time (cpimh_model_*) with injected noise."""
    rng = np.random.default_rng(8)
    for name, c in model_cases(up, oracle, golden).items():
        N = 300
        ocfg = oracle.make_config(c["oid"], c["d"], c["dy"], N, 10, theta=c["otheta"], **c["okw"])
        ni, nt = oracle.n_init(ocfg), oracle.n_trans(ocfg)
        if name == "levy_sv":
            zi = np.stack([rng.gamma(7.0, 0.11, N), rng.exponential(0.02, N), rng.exponential(0.03, N)])
            zt = np.stack([rng.exponential(0.02, N), rng.exponential(0.03, N)])
        elif name == "sk":
            zi = None
            zt = np.concatenate([rng.exponential(size=(nt // 2, N)), rng.random((nt // 2, N))])
        else:
            zi, zt = rng.normal(size=(ni, N)), rng.normal(size=(nt, N))
        x0g = c["model"].rinit(N, c["theta"], zi)
        x0o = oracle.model_rinit(ocfg, zi if zi is not None else np.zeros((1, N)))
        assert np.allclose(x0g, x0o, rtol=1e-13, atol=1e-14), name
        x1g = c["model"].rtransition(x0o, c["theta"], 3, zt, sk_max_events=c["okw"].get("sk_max_events", 64))
        x1o = oracle.model_rtransition(ocfg, x0o, 3, zt)
        assert np.allclose(x1g, x1o, rtol=1e-11, atol=1e-12), name
        yv = c["obs"][2]
        if name == "levy_sv":
            x1o = np.abs(x1o) + 1e-3
        assert np.allclose(c["model"].dmeasurement(x1o, c["theta"], yv), oracle.model_dmeasurement(ocfg, x1o, yv), rtol=1e-12, atol=1e-12), name
    x = rng.normal(size=(64, 50)) + 2
    got = up.lorenz_transition(x, 0.0, 0.05, 0.05, [8.0], integrator="rk4", substeps=3)            # warp-per-particle path (d = 64)
    ref = np.stack([oracle.lorenz_step(x[:, n], 8.0, 0, 3) for n in range(50)], axis=1)
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-13)
    x8 = rng.normal(size=(8, 50)) + 2
    got = up.lorenz_transition(x8, 0.1, 0.15, 0.05, [8.0])                                          # dopri5, the reference's integrator
    ref = np.stack([oracle.lorenz_step(x8[:, n], 8.0, 1, 1, 0.1, 0.15, 0.05) for n in range(50)], axis=1)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("N", [1, 2, 10, 31, 33, 110])
@pytest.mark.parametrize("store_paths", [2, 3])
def test_small_and_ragged_particle_counts(up, oracle, golden, N, store_paths):
    """N from the reference's own experiments (10..110, inst/reproduce/lgssm/lgssm_settings.R:8) down to a single particle."""
    y = golden["lgssm_y"][:30]
    th = [0.5, math.sqrt(10.0)]
    r = up.cpf_batch(N, up.get_ar(1), th, y, 5, seed=77, logz_t=True, store_paths=store_paths)
    ocfg = oracle.make_config(oracle.MODEL_AR, 1, 1, N, 30, theta=th)
    for b in range(5):
        o = oracle.cpf(ocfg, y, seed=77, filter_id=b)
        assert np.allclose(r["logz_t"][b], o["logz_t"], rtol=RTOL_LOGZ, atol=0)
        assert int(r["path_index"][b]) == o["path_index"] and np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)
    one = up.cpf_batch(N, up.get_ar(1), th, y[:1], 3, seed=5, logz_t=True, store_paths=store_paths)        # T = 1
    assert one["trajectory"].shape == (3, 1, 2) and np.all(np.isfinite(one["logz"]))


def test_multi_wave_batches_and_tail_launch(up, golden):
    """More filters than resident CTAs: whole waves + a 1024-thread tail launch must agree with the same filters run in
    pieces.  The block scans/reductions associate differently for different CTA sizes, so filters that move between the
    512- and 1024-thread variants agree to rounding (1e-12 relative), not bit for bit; same geometry => same bits."""
    th = golden["sv_theta_mle"]
    y = golden["sv_y"][:12]
    N = 2048
    pb = up.PfBatch(up.get_levy_sv(), th, y, N, 700)
    eff = pb.effective_config()
    resident = 148 * eff["ctas_per_sm"]
    B = resident + 57 if resident + 57 <= 700 else 700
    pb.run(B, 21)
    a = pb.fetch(path_index=True)
    pb.run(100, 21, first_filter=B - 100)
    b = pb.fetch(path_index=True)
    assert np.allclose(a["logz"][B - 100:], b["logz"], rtol=1e-12, atol=0) and np.allclose(a["trajectory"][B - 100:], b["trajectory"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(a["path_index"][B - 100:], b["path_index"])
    assert pb.launch_count == (3 if B > resident else 2)
    pb.run(B, 21)
    c = pb.fetch(path_index=True)
    assert np.array_equal(a["logz"], c["logz"]) and np.array_equal(a["trajectory"], c["trajectory"])
    pb.close()


@pytest.mark.parametrize("d", [6, 7, 9, 11, 12, 13, 14, 15])
def test_get_ar_every_dimension_up_to_sixteen(up, oracle, d):
    """R/model_get_ar.R:11-46 accepts any dimension: every d <= 16 has a kernel (round 1: only 1,2,3,4,5,8,10,16).  Injected
    noise: identical ancestors and log Z; Philox seed for seed."""
    rng = np.random.default_rng(d)
    N, T, B = 96, 12, 2
    th = [0.4, 1.3]
    obs = rng.normal(size=(T, d)) * 1.5
    ocfg = oracle.make_config(oracle.MODEL_AR, d, d, N, T, theta=th)
    nz = injected_noise(rng, oracle, ocfg, B, N, T, 0)
    pb = up.PfBatch(up.get_ar(d), th, obs, N, B)
    pb.record_ancestors(True); pb.set_noise(B, nz); pb.run(B, 0)
    r = pb.fetch(logz_t=True, path_index=True)
    for b in range(B):
        o = oracle.cpf(ocfg, obs, noise={k: v[b] for k, v in nz.items()}, want_ancestors=True)
        assert np.array_equal(pb.fetch_ancestors(b), o["ancestors"])
        assert np.max(np.abs(r["logz_t"][b] - o["logz_t"]) / np.abs(o["logz_t"])) < RTOL_LOGZ
        assert np.allclose(r["trajectory"][b], o["trajectory"], rtol=1e-9, atol=1e-9)
    pb.close()
    r = up.cpf_batch(N, up.get_ar(d), th, obs, 2, seed=5, logz_t=True)
    for b in range(2):
        o = oracle.cpf(ocfg, obs, seed=5, filter_id=b)
        assert np.max(np.abs(r["logz_t"][b] - o["logz_t"]) / np.abs(o["logz_t"])) < RTOL_LOGZ and int(r["path_index"][b]) == o["path_index"]


def test_one_shot_entries_cache_their_workspaces_and_match_the_handle_api(up, golden):
    """cpimh_pf_batch / cpimh_pf_batch_all_particles (what the R shim calls) keep their HBM workspaces between calls: repeated
    calls, calls with another configuration in between and calls after cpimh_release_cached() all return the handle API's bits."""
    th, y = [0.5, math.sqrt(10.0)], golden["lgssm_y"][:30]
    pb = up.PfBatch(up.get_ar(1), th, y, 128, 6); pb.all_particles(True); pb.run(6, 9, first_filter=2)
    ref = pb.fetch(); ref_exp = pb.fetch_exp_paths(6); pb.close()
    for rep in range(3):
        r = up.cpf_batch(128, up.get_ar(1), th, y, 6, seed=9, first_filter=2)
        assert np.array_equal(r["logz"], ref["logz"]) and np.array_equal(r["trajectory"], ref["trajectory"])
        if rep == 0:
            up.cpf_batch(64, up.get_gss(), [], golden["gss_y"][:10], 3, seed=1)          # evicts the cached handle
        if rep == 1:
            up.lib().cpimh_release_cached()
    ra = up.cpf_batch(128, up.get_ar(1), th, y, 6, seed=9, first_filter=2, all_particles=True)
    assert np.array_equal(ra["logz"], ref["logz"]) and np.array_equal(ra["exp_path"], ref_exp)
    r = up.cpf_batch(128, up.get_ar(1), th, y, 4, seed=9, first_filter=2)                  # smaller batch reuses the larger handle
    assert np.array_equal(r["logz"], ref["logz"][:4])


def test_chain_iteration_index_has_twenty_bits(up, oracle, golden):
    """The iteration of a chain is the high word of the 64-bit filter id; 20 bits of it reach the Philox counter (round 1: 12, so
    iteration 4096 + x replayed the draws of iteration x).  CPU and GPU agree seed for seed above 4095, and nothing aliases."""
    th, y = [0.5, math.sqrt(10.0)], golden["lgssm_y"][:15]
    ocfg = oracle.make_config(oracle.MODEL_AR, 1, 1, 64, 15, theta=th)
    lz = {}
    for it in (5, 4096 + 5, 70000, 0xFFFFF):
        fid = 7 | (it << 32)
        r = up.cpf_batch(64, up.get_ar(1), th, y, 1, seed=3, first_filter=fid, logz_t=True)
        o = oracle.cpf(ocfg, y, seed=3, filter_id=fid)
        assert np.max(np.abs(r["logz_t"][0] - o["logz_t"]) / np.abs(o["logz_t"])) < RTOL_LOGZ and int(r["path_index"][0]) == o["path_index"]
        lz[it] = r["logz"][0]
    assert len(set(lz.values())) == 4


def test_levy_sv_observation_no_particle_explains(up, oracle, golden):
    """levy_sv forms its weights as exp(a) v^(-1/2) with a <= 0 and no maximum subtracted; when every a is below -600 (an absurd
    observation) the kernel falls back to the maximum-relative form: log Z stays finite and equal to the oracle's."""
    th = golden["sv_theta_mle"]
    y = golden["sv_y"][:6].copy()
    y[3, 0] = 250.0                                   # (y - mu - beta v)^2 / (2 v) ~ 1e4 for every particle
    r = up.cpf_batch(512, up.get_levy_sv(), th, y, 2, seed=8, logz_t=True)
    ocfg = oracle.make_config(oracle.MODEL_LEVY_SV, 2, 1, 512, 6, theta=th)
    for b in range(2):
        o = oracle.cpf(ocfg, y, seed=8, filter_id=b)
        assert np.all(np.isfinite(r["logz_t"][b])) and o["logz_t"][3] < -600
        assert np.max(np.abs(r["logz_t"][b] - o["logz_t"]) / np.abs(o["logz_t"])) < RTOL_LOGZ and int(r["path_index"][b]) == o["path_index"]


@pytest.mark.parametrize("name,N,extra", [("gss", 256, 100), ("gss", 256, 0), ("levy_sv", 2048, 5), ("ar3", 512, 301)])
def test_one_shot_entry_over_several_waves_is_pipelined_and_matches_the_handle_api(up, golden, name, N, extra):
    """More filters than resident CTAs: the one-shot entry runs groups of whole waves (+ the tail launch) while it copies the
    finished groups' results to the caller; every output must equal the handle API's run of the same batch, bit for bit."""
    model, theta, obs = {
        "gss": (up.get_gss(), [], golden["gss_y"][:100]),          # 1.6 KB of results per filter: 7 waves are copied in two groups
        "levy_sv": (up.get_levy_sv(), golden["sv_theta_mle"], golden["sv_y"][:6]),
        "ar3": (up.get_ar(3), [0.4, 1.0], np.random.default_rng(3).normal(size=(8, 3))),
    }[name]
    pb = up.PfBatch(model, theta, obs, N, 8)
    grid = pb.effective_config()["ctas_per_sm"] * 148
    pb.close()
    B = (7 if name == "gss" else 2) * grid + extra
    pb = up.PfBatch(model, theta, obs, N, B)
    pb.run(B, 21, first_filter=4)
    ref = pb.fetch(logz_t=True, path_index=True)
    pb.close()
    r = up.cpf_batch(N, model, theta, obs, B, seed=21, first_filter=4, logz_t=True)
    for k in ("logz", "logz_t", "trajectory", "path_index"):
        assert np.array_equal(r[k], ref[k]), k


@pytest.mark.parametrize("name,N", [("levy_sv", 4096), ("gss", 2000), ("ar1", 1025), ("ar1", 37)])
def test_shuffled_ancestors_philox_parity(up, oracle, golden, name, N):
    """cfg.shuffle = 1 (std::random_shuffle of the ancestors, src/multinomial.cpp:6-10,29-30) with Philox uniforms: the
    kernel replays the N - 1 swaps in parallel (every final position traced back to its origin); ancestors must equal the
    oracle's one-swap-after-the-other loop at every step."""
    c = model_cases(up, oracle, golden)[name]
    obs = c["obs"][:6]
    B, seed = 2, 77
    pb = up.PfBatch(c["model"], c["theta"], obs, N, B, shuffle=1)
    pb.record_ancestors(True)
    pb.run(B, seed, first_filter=1)
    r = pb.fetch(logz_t=True, path_index=True)
    okw = {k: v for k, v in c["okw"].items() if k != "sk_max_events"}
    ocfg = oracle.make_config(c["oid"], c["d"], c["dy"], N, obs.shape[0], theta=c["otheta"], shuffle=1, **okw)
    for b in range(B):
        o = oracle.cpf(ocfg, obs, seed=seed, filter_id=1 + b, want_ancestors=True)
        anc = pb.fetch_ancestors(b)
        assert np.array_equal(anc, o["ancestors"]), (name, b, int(np.sum(anc != o["ancestors"])))
        assert not np.all(np.diff(anc[1]) >= 0)                                   # really shuffled
        assert int(r["path_index"][b]) == o["path_index"]
    pb.close()


@pytest.mark.parametrize("kind", ["all_to_front", "shift_by_one", "identity", "random"])
def test_shuffle_replay_adversarial_draws(up, oracle, golden, kind):
    """Injected shuffle uniforms that stress the parallel replay: every swap with position 0 (one bucket of N - 1 entries: the
    one-thread fallback), j_i = i - 1 (a hop chain of length N from the last position), j_i = i (no swap at all)."""
    N, T, B = 1500, 3, 2                     # >= 1024: the parallel replay (smaller filters swap serially)
    c = model_cases(up, oracle, golden)["ar1"]
    obs = c["obs"][:T]
    rng = np.random.default_rng(5)
    ocfg = oracle.make_config(c["oid"], c["d"], c["dy"], N, T, theta=c["otheta"], shuffle=1, **c["okw"])
    nz = injected_noise(rng, oracle, ocfg, B, N, T, 1)
    i = np.arange(1, N)
    u = {"all_to_front": np.zeros(N - 1), "shift_by_one": (i - 0.5) / (i + 1), "identity": (i + 0.5) / (i + 1), "random": rng.random(N - 1)}[kind]
    nz["shuffle"] = np.broadcast_to(u, (B, T, N - 1)).copy()
    pb = up.PfBatch(c["model"], c["theta"], obs, N, B, shuffle=1)
    pb.record_ancestors(True)
    pb.set_noise(B, nz)
    pb.run(B, 0)
    pb.fetch()
    for b in range(B):
        o = oracle.cpf(ocfg, obs, noise={k: v[b] for k, v in nz.items()}, want_ancestors=True)
        assert np.array_equal(pb.fetch_ancestors(b), o["ancestors"]), (kind, b)
    pb.close()
