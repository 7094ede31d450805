#!/usr/bin/env python
"""Driver for profiling the wide (Lorenz-96 d=64) filter: N=65536, short horizon."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up
import torch
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
rng = np.random.default_rng(17)
y = rng.normal(size=(T, 64)) * 1.5 + 2.0
pb = up.PfBatch(up.get_lorenz(1.0, 64, "rk4", 1), [8.0, 0.5], y, N, B)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    pb.run(B, 1 + rep); pb.fetch(trajectory=False)
    dt = time.time() - t0
    print(f"lorenz64 N={N} T={T} B={B}: {dt*1e3:.2f} ms  {N*T*B/dt:.3e} particle-steps/s  {N*T*B/dt*1568/1e9:.0f} GB/s alg")
