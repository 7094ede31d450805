#!/usr/bin/env python
"""cfg.shuffle = 1 (the reference's std::random_shuffle of the ancestors, src/multinomial.cpp:29-30) against the default 0:
device time of config 2's shape.  python tools/shuffle_timing.py [T B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unbiased_pimh_b200 as up
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 444
for name, m, th, y, N in (("levy_sv", up.get_levy_sv(), G["sv_theta_mle"], np.tile(G["sv_y"], (3, 1))[:T], 4096),
                          ("gss", up.get_gss(), [], np.tile(G["gss_y"], (12, 1))[:T], 256)):
    for sh in (0, 1):
        pb = up.PfBatch(m, th, y, N, B, shuffle=sh)
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
        print("%s N=%d T=%d B=%d shuffle=%d (%s): %8.2f ms  %.3e particle-steps/s  mean logz %.4f" % (name, N, T, B, sh, pb.effective_config(), ms, N * T * B / ms * 1e3, r["logz"].mean()))
        pb.close()
