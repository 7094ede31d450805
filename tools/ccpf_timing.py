#!/usr/bin/env python
"""Wall time of the one-shot coupled-CCPF chain entry (cpimh_coupled_ccpf), repeated calls."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import unbiased_pimh_b200 as up
from bench import simulate_ar
obs1 = simulate_ar(100, 1, 0.5, float(np.sqrt(10.0)))
m1, th1 = up.get_ar(1), [0.5, float(np.sqrt(10.0))]
up.coupled_ccpf_chains_batch(m1, th1, obs1, 128, 2000, K=10, k=3, with_as=True, seed=5)
for rep in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = up.coupled_ccpf_chains_batch(m1, th1, obs1, 128, 10000, K=10, k=3, with_as=True, seed=6 + rep)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms  %.1f k replicates/s  filters %d  mean tau %.3f" % (rep, dt * 1e3, 10000 / dt / 1e3, g["nfilters"], g["meetingtime"].mean()))
