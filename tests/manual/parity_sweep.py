"""Ad-hoc parity sweep over ragged particle counts: bootstrap / conditional / adaptive filters on the GPU against the CPU
oracle on the same Philox streams (ancestors identical, log Z within 1e-9).  Test infrastructure, like tests/."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
y = G["lgssm_y"][:25]
th = [0.5, math.sqrt(10.0)]
bad = 0
sizes = list(range(2, 75)) + [95, 96, 97, 127, 128, 129, 255, 257, 511, 513, 1000, 1023, 1025, 2047, 2049, 3000]
for N in sizes:
    cfg = O.make_config(O.MODEL_AR, 1, 1, N, 25, theta=th)
    for sp in (2, 3):
        pb = up.PfBatch(up.get_ar(1), th, y, N, 2, store_paths=sp)
        pb.record_ancestors(True); pb.run(2, 7, first_filter=N); r = pb.fetch(logz_t=True)
        for b in range(2):
            o = O.cpf(cfg, y, seed=7, filter_id=N + b, want_ancestors=True)
            if not np.array_equal(pb.fetch_ancestors(b), o["ancestors"]) or abs(r["logz"][b] - o["logz"]) > 1e-9 * abs(o["logz"]):
                bad += 1; print("cpf mismatch N=%d sp=%d b=%d" % (N, sp, b))
        pb.close()
    refs = [O.cpf(cfg, y, seed=9, filter_id=i)["trajectory"] for i in range(2)]
    g = up.ccpf_batch(N, up.get_ar(1), th, y, refs[:1], refs[1:], with_as=True, seed=11, first_filter=N, ancestors=True)
    o = O.ccpf(cfg, y, refs[0], refs[1], with_as=True, seed=11, filter_id=N, want_ancestors=True)
    if not (np.array_equal(g["ancestors"][0, 0], o["ancestors1"]) and np.array_equal(g["ancestors"][0, 1], o["ancestors2"])):
        bad += 1; print("ccpf mismatch N=%d" % N)
    for systematic in (False, True):
        a = up.pf_batch(y, th, up.get_ar(1), N, 1, ess_threshold=0.6, systematic=systematic, seed=13, first_filter=N, ancestors=True)
        o = O.pf_adaptive(cfg, y, seed=13, filter_id=N, ess_threshold=0.6, systematic=systematic, want_ancestors=True)
        if not np.array_equal(a["ancestors"][0], o["ancestors"]) or abs(a["loglik"][0] - o["loglik"]) > 1e-9 * abs(o["loglik"]):
            bad += 1; print("pf mismatch N=%d systematic=%s" % (N, systematic))
print("parity sweep over %d sizes: %d mismatches" % (len(sizes), bad))
