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
