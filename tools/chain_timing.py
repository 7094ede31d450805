#!/usr/bin/env python
"""Device-time breakdown of the coupled-PIMH driver (one GPU): CUDA events around cpimh_chains_run for several batch sizes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import unbiased_pimh_b200 as up
from bench import simulate_ar
obs = simulate_ar(100, 10, 0.5, float(np.sqrt(10.0)))
m, th = up.get_ar(10), [0.5, float(np.sqrt(10.0))]
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for R in (5000, 20000, 50000):
    cb = up.ChainBatch(m, th, obs, 1024, R)
    st = torch.cuda.current_stream().cuda_stream
    cb.run(R, 3, first, 0, 1, -1, stream=st); cb.fetch(hbar=False)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(); cb.run(R, 4 + rep, first, 0, 1, -1, stream=st); e1.record()
        r = cb.fetch(hbar=False)
        wall = time.perf_counter() - t0
        print("R=%6d first=%d: device %.1f ms, wall incl. fetch %.1f ms, %.1f k replicates/s, filters needed %d, run %d, max tau %d" % (
            R, first, e0.elapsed_time(e1), wall * 1e3, R / wall / 1e3, r["nfilters"], up.lib().cpimh_chains_filters_launched(cb.h), r["meetingtime"].max()))
    cb.close()
