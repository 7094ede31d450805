#!/usr/bin/env python
"""Driver for profiling the coupled-PIMH persistent kernel (config-4 shape: get_ar(10), N=1024, T=100)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up
import torch
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rng = np.random.default_rng(17)
A = np.array([[0.5 ** (1 + abs(i - j)) for j in range(d)] for i in range(d)])
x = rng.normal(size=d); y = np.empty((100, d))
for t in range(100):
    x = A @ x + rng.normal(size=d); y[t] = x + np.sqrt(10.0) * rng.normal(size=d)
cb = up.ChainBatch(up.get_ar(d), [0.5, float(np.sqrt(10.0))], y, N, R)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    cb.run(R, 5 + rep); out = cb.fetch(hbar=False)
    dt = time.time() - t0
    print(f"chains ar{d} N={N} R={R}: {dt*1e3:.1f} ms  {R/dt:.0f} replicates/s  filters {out['nfilters']}  {out['nfilters']*100*N/dt:.3e} particle-steps/s  mean tau {out['meetingtime'].mean():.3f}")
