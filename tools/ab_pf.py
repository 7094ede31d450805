#!/usr/bin/env python
"""A/B timing of the fused filter kernel with CUDA events (device time, HBM-resident), for a list of library builds:
    python tools/ab_pf.py model N T B [lib1.so lib2.so ...]
Each library is timed in its own process (CPIMH_LIB selects it)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("AB_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import unbiased_pimh_b200 as up
    G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
    model, N, T, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    if model == "levy_sv":
        m, th, y = up.get_levy_sv(), G["sv_theta_mle"], np.tile(G["sv_y"], (3, 1))[:T]
    elif model == "ar10":
        m, th, y = up.get_ar(10), [0.5, np.sqrt(10)], np.random.default_rng(1).normal(size=(T, 10))
    elif model == "sk":
        A = np.array([[1, 0, 0, 0], [0, 1, 2, 0]], dtype=float)
        m, th, y = up.get_stochastic_kinetic(), {"A_obs": A, "s2_obs": 1.0}, np.tile(G["sk_y"], (3, 1))[:T]
    else:
        m, th, y = up.get_gss(), [], np.tile(G["gss_y"], (12, 1))[:T]
    pb = up.PfBatch(m, th, y, N, B)
    for rep in range(2):
        pb.run(B, 1 + rep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(3):
        pb.run(B, 10 + rep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    r = pb.fetch(trajectory=False)
    pb.close()
    print("%-40s %s N=%d T=%d B=%d: %8.2f ms  %.3e particle-steps/s  mean logz %.6f" % (os.path.basename(os.environ.get("CPIMH_LIB", "libcpimh.so")), model, N, T, B, ms, N * T * B / ms * 1e3, r["logz"].mean()))
else:
    libs = sys.argv[5:] or [os.path.join(ROOT, "unbiased_pimh_b200", "libcpimh.so")]
    for lib in libs:
        env = dict(os.environ, AB_CHILD="1", CPIMH_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:5], env=env)
