import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import unbiased_pimh_b200 as up
T, N, B = 100, 1024, 16384
y = np.random.default_rng(1).normal(size=(T, 10))
for nt in (0, 128, 256, 512):
    pb = up.PfBatch(up.get_ar(10), [0.5, np.sqrt(10)], y, N, B, threads_per_filter=nt)
    for rep in range(2): pb.run(B, 1 + rep)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(3): pb.run(B, 10 + rep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("ar10 nt=%d %s: %.2f ms %.3e" % (nt, pb.effective_config(), ms, N * T * B / ms * 1e3))
    pb.close()
