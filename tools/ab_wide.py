#!/usr/bin/env python
"""A/B timing of the wide (Lorenz-96 d=64, N=65536) filter for a list of library builds: python tools/ab_wide.py T lib1.so lib2.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("AB_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import unbiased_pimh_b200 as up
    T = int(sys.argv[1]); N = 65536
    y = np.random.default_rng(17).normal(size=(T, 64)) * 1.5 + 2.0
    pb = up.PfBatch(up.get_lorenz(1.0, 64, "rk4", 1), [8.0, 0.5], y, N, 1)
    for rep in range(2):
        pb.run(1, 1 + rep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(3):
        pb.run(1, 10 + rep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    r = pb.fetch(trajectory=False); pb.close()
    print("%-30s lorenz64 N=%d T=%d: %8.2f ms  %.3e particle-steps/s  logz %.6f" % (os.path.basename(os.environ.get("CPIMH_LIB", "libcpimh.so")), N, T, ms, N * T / ms * 1e3, r["logz"][0]))
else:
    for lib in sys.argv[2:]:
        subprocess.run([sys.executable, os.path.abspath(__file__), sys.argv[1]], env=dict(os.environ, AB_CHILD="1", CPIMH_LIB=os.path.abspath(lib)))
