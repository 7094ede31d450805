#!/usr/bin/env python
"""Small driver for ncu: one launch of the fused filter kernel on a config-2-shaped workload (levy_sv, N=4096)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
model = sys.argv[1] if len(sys.argv) > 1 else "levy_sv"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 50
B = int(sys.argv[4]) if len(sys.argv) > 4 else 296
cap = int(sys.argv[5]) if len(sys.argv) > 5 else 0
nt = int(sys.argv[6]) if len(sys.argv) > 6 else 0
if model == "levy_sv":
    m, th, y = up.get_levy_sv(), G["sv_theta_mle"], np.tile(G["sv_y"], (3, 1))[:T]
elif model == "ar10":
    m, th, y = up.get_ar(10), [0.5, np.sqrt(10)], np.random.default_rng(1).normal(size=(T, 10))
elif model == "sk":
    A = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
    m, th, y = up.get_stochastic_kinetic(), {"A_obs": A, "s2_obs": 1.0}, np.tile(G["sk_y"], (3, 1))[:T]
else:
    m, th, y = up.get_gss(), [], np.tile(G["gss_y"], (12, 1))[:T]
pb = up.PfBatch(m, th, y, N, B, tree_capacity=cap, threads_per_filter=nt)
print(pb.effective_config())
import time
import torch
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    pb.run(B, 1 + rep)
    pb.fetch(trajectory=False)
    dt = time.time() - t0
    print(f"{model} N={N} T={T} B={B}: {dt*1e3:.2f} ms  {N*T*B/dt:.3e} particle-steps/s")
