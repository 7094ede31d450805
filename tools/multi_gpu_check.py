#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/multi_gpu_check.py : coupled-PIMH unbiased estimator with replicates sharded over G
GPUs and ONE NCCL all-reduce; rank 0 checks the result against a single-GPU run of the same global replicate ids."""
import math
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
g = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
model, theta, y = up.get_ar(1), [0.5, math.sqrt(10.0)], g["lgssm_y"]
R = 20000
dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
res = up.unbiased_estimator_sharded(model, theta, y, 128, R, K=2, k=1, seed=5)
dist.barrier(); torch.cuda.synchronize(); dt = time.time() - t0
if rank == 0:
    full = up.coupled_pimh_chains_batch(model, theta, y, 128, R, K=2, k=1, seed=5)
    ref = full["hbar"][:, 0, :].mean(axis=0)
    err = float(np.max(np.abs(res["mean"][0] - ref)))
    print(f"world={world} replicates={res['nreplicates']} filters={res['nfilters']} (single GPU {full['nfilters']}) "
          f"max|mean diff|={err:.3e} time={dt:.3f}s -> {R/dt:.0f} replicates/s")
    assert res["nreplicates"] == R and res["nfilters"] == full["nfilters"] and err < 1e-12
    print("MULTI_GPU_OK")
dist.destroy_process_group()
