#!/usr/bin/env python
"""First-light diagnostics on a B200: seed-for-seed comparison of the CUDA filter with the CPU oracle (Philox streams
shared), injected-noise parity, and a quick throughput probe.  Scratch tool, not part of the test-suite."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))


def compare(name, model, theta, otheta, omodel, obs, N, B=4, seed=17, **okw):
    T = obs.shape[0]
    try:
        r = up.cpf_batch(N, model, theta, obs, B, seed=seed, logz_t=True)
    except Exception as e:
        print(f"[{name}] GPU FAILED: {e}")
        traceback.print_exc()
        return
    cfg = O.make_config(omodel, model.dimension, model.dimension_y, N, T, theta=otheta, **okw)
    worst = 0.0; wtraj = 0.0; pidx_ok = True
    for b in range(B):
        o = O.cpf(cfg, obs, seed=seed, filter_id=b)
        rel = np.max(np.abs(r["logz_t"][b] - o["logz_t"]) / np.maximum(np.abs(o["logz_t"]), 1e-300))
        worst = max(worst, rel)
        wtraj = max(wtraj, np.max(np.abs(r["trajectory"][b] - o["trajectory"])))
        pidx_ok &= (int(r["path_index"][b]) == o["path_index"])
    print(f"[{name}] N={N} T={T} B={B}: max rel |dlogz_t| = {worst:.3e}, max |dtraj| = {wtraj:.3e}, path index equal: {pidx_ok}, logz[0] = {r['logz'][0]:.10f}")


def main():
    import torch
    print("device:", torch.cuda.get_device_name(0))
    compare("gss", up.get_gss(), [], [], O.MODEL_GSS, G["gss_y"], 256)
    compare("gss", up.get_gss(), [], [], O.MODEL_GSS, G["gss_y"], 1000)
    compare("ar1", up.get_ar(1), [0.5, np.sqrt(10)], [0.5, np.sqrt(10)], O.MODEL_AR, G["lgssm_y"], 150)
    rng = np.random.default_rng(5)
    y10 = rng.normal(size=(50, 10)) * 2
    compare("ar10", up.get_ar(10), [0.5, np.sqrt(10)], [0.5, np.sqrt(10)], O.MODEL_AR, y10, 1024)
    th = G["sv_theta_mle"]
    compare("levy_sv", up.get_levy_sv(), th, th, O.MODEL_LEVY_SV, G["sv_y"][:200], 4096, B=2)
    A = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
    compare("sk", up.get_stochastic_kinetic(), {"A_obs": A, "s2_obs": 1.0}, [1.0] + list(A.ravel(order="F")), O.MODEL_SK, G["sk_y"], 2048, B=2)
    yl = rng.normal(size=(20, 8)) * 3 + 2
    compare("lorenz8-rk4", up.get_lorenz(1.0, 8, "rk4", 2), [8.0, 0.5], [8.0, 0.5, 1.0], O.MODEL_LORENZ, yl, 512, B=2, integrator=0, substeps=2)
    compare("lorenz8-dopri5", up.get_lorenz(1.0, 8, "dopri5"), [8.0, 0.5], [8.0, 0.5, 1.0], O.MODEL_LORENZ, yl, 512, B=2, integrator=1)

    # throughput probe: config 2 shape at reduced batch
    for (N, T, B) in ((4096, 200, 296), (4096, 1000, 1024)):
        y = np.tile(G["sv_y"], (2, 1))[:T]
        pb = up.PfBatch(up.get_levy_sv(), th, y, N, B)
        print("effective:", pb.effective_config(), "HBM MB:", pb.device_bytes / 1e6)
        pb.run(B, 1); pb.fetch(trajectory=False)
        torch.cuda.synchronize()
        t0 = time.time(); pb.run(B, 2); pb.fetch(trajectory=False); dt = time.time() - t0
        print(f"levy_sv N={N} T={T} B={B}: {dt*1e3:.1f} ms -> {N*T*B/dt:.3e} particle-steps/s ({N*T*B/dt*64/1e9:.1f} GB/s algorithmic)")
        pb.close()
    # coupled chains smoke
    r = up.coupled_pimh_chains_batch(up.get_ar(1), [0.5, np.sqrt(10)], G["lgssm_y"], 100, 2000, K=1, seed=3)
    mt = r["meetingtime"]
    print("coupled PIMH ar1 N=100: mean tau", mt.mean(), "P[tau=1]", (mt == 1).mean(), "filters", r["nfilters"], "hbar[0,:3]", r["hbar"][0, 0, :3])
    cfg = O.make_config(O.MODEL_AR, 1, 1, 100, 100, theta=[0.5, np.sqrt(10)])
    o = O.coupled_pimh_batch(cfg, G["lgssm_y"], 3, 0, 50)
    print("oracle vs gpu meeting times equal (first 50):", np.array_equal(np.where(o["meetingtime"] < 0, -1, o["meetingtime"]), mt[:50]),
          "max |dhbar|", np.max(np.abs(o["hbar"] - r["hbar"][:50, 0, :])))


if __name__ == "__main__":
    main()
