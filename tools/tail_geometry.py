#!/usr/bin/env python
"""Timing of a PARTIAL wave of the fused filter (the tail of config 2: 136 filters on 148 SMs) for every CTA size:
    python tools/tail_geometry.py [model N T B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unbiased_pimh_b200 as up
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
B = int(sys.argv[4]) if len(sys.argv) > 4 else 136
m, th, y = up.get_levy_sv(), G["sv_theta_mle"], np.tile(G["sv_y"], (3, 1))[:T]
for nt in (128, 256, 512, 1024):
    try:
        pb = up.PfBatch(m, th, y, N, B, threads_per_filter=nt)
    except Exception as e:
        print(nt, "unavailable:", str(e)[:80]); continue
    for rep in range(2):
        pb.run(B, 1 + rep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(3):
        pb.run(B, 10 + rep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("levy_sv N=%d T=%d B=%d threads=%4d (%s): %7.2f ms" % (N, T, B, nt, pb.effective_config(), ms))
    pb.close()
