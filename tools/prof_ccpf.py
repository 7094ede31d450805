#!/usr/bin/env python
"""Timing driver for the conditional filters: cpimh_ccpf_batch (one launch) and cpimh_coupled_ccpf (chains)."""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
what = sys.argv[1] if len(sys.argv) > 1 else "all"
TH = [0.5, math.sqrt(10.0)]


def batch(model, th, y, N, B, nsys, with_as, reps=3):
    T = y.shape[0]
    x = up.cpf_batch(N, model, th, y, 2 * B, seed=1)["trajectory"]
    best = 1e9
    for r in range(reps):
        t0 = time.time()
        up.ccpf_batch(N, model, th, y, x[:B], x[B:] if nsys == 2 else None, with_as=with_as, seed=2 + r)
        best = min(best, time.time() - t0)
    ps = nsys * N * T * B
    print(f"ccpf_batch d={model.dimension} N={N} T={T} B={B} nsys={nsys} as={with_as}: {best*1e3:.1f} ms (incl. alloc + H2D/D2H)  {ps/best:.3e} particle-steps/s")


def chains(model, th, y, N, R, K, with_as, max_it=-1):
    t0 = time.time()
    g = up.coupled_ccpf_chains_batch(model, th, y, N, R, K=K, k=0, with_as=with_as, seed=5, max_iterations=max_it)
    dt = time.time() - t0
    T = y.shape[0]
    print(f"coupled_ccpf d={model.dimension} N={N} T={T} R={R} K={K} as={with_as}: {dt:.3f} s  {R/dt:.0f} replicates/s  filters {g['nfilters']}  "
          f"mean tau {g['meetingtime'].mean():.2f} max {g['meetingtime'].max()}  {g['nfilters']*N*T/dt:.3e} filter-particle-steps/s")


if what in ("all", "batch"):
    y = G["lgssm_y"]
    batch(up.get_ar(1), TH, y, 128, 4000, 2, True)
    batch(up.get_ar(1), TH, y, 1024, 2000, 2, False)
    batch(up.get_ar(1), TH, y, 1024, 2000, 1, False)
    y10 = np.random.default_rng(1).normal(size=(100, 10))
    batch(up.get_ar(10), TH, y10, 1024, 1000, 2, True)
    batch(up.get_levy_sv(), G["sv_theta_mle"], np.tile(G["sv_y"], (2, 1))[:200], 4096, 296, 2, False)
if what in ("all", "chains"):
    y = G["lgssm_y"]
    chains(up.get_ar(1), TH, y, 128, 20000, 10, True)
    chains(up.get_ar(1), TH, y, 128, 5000, 10, False, max_it=1000)
    y10 = np.random.default_rng(1).normal(size=(100, 10))
    chains(up.get_ar(10), TH, y10, 1024, 4000, 5, True)
