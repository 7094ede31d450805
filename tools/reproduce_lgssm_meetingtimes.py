#!/usr/bin/env python
"""The reference's headline experiment on one GPU: inst/reproduce/lgssm/lgssm_get_and_plot_meetingtimes.R
("around 30 mins on a 90 core server", :15): get_ar(1), theta = (0.5, sqrt(10)), the shipped data100obs data set, for
N in {10, 30, 50, 70, 90, 110}: 10^4 bootstrap filters for Var(log Z) and 10^5 coupled-PIMH replicates (K = 1) for the
meeting-time distribution, compared with the closed-form law P[tau = n] of the same script (:210-231).
Prints one line per N and the total wall time."""
import math
import os
import sys
import time

import numpy as np
from scipy.stats import norm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unbiased_pimh_b200 as up  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))
y, th, model = g["lgssm_y"], [0.5, math.sqrt(10.0)], up.get_ar(1)
n_var, R = 10000, 100000
t_all = time.time()
for N in (10, 30, 50, 70, 90, 110):
    t0 = time.time()
    lz = up.cpf_batch(N, model, th, y, n_var, seed=1000 + N, store_paths=0)["logz"]
    s2 = float(lz.var(ddof=1)); s = math.sqrt(s2)
    r = up.coupled_pimh_chains_batch(model, th, y, N, R, K=1, seed=2000 + N)
    mt = r["meetingtime"]
    z = np.linspace(-12 * s, 12 * s, 40001)
    p = norm.cdf(-z / s) + np.exp(-z + s2 / 2) * norm.cdf((z - s2) / s)
    pred = [float(np.sum(p * (1 - p) ** (n - 1) * norm.pdf(z, scale=s)) * (z[1] - z[0])) for n in (1, 2, 3)]
    emp = [float(np.mean(mt == n)) for n in (1, 2, 3)]
    print(f"N={N:4d}  Var(logZ)={s2:.3f}  mean tau={mt.mean():.3f}  P[tau=1,2,3] empirical {emp[0]:.4f} {emp[1]:.4f} {emp[2]:.4f}"
          f" | closed form {pred[0]:.4f} {pred[1]:.4f} {pred[2]:.4f}  filters={n_var + r['nfilters']}  {time.time() - t0:.2f} s")
print(f"total wall time {time.time() - t_all:.2f} s (reference: 'around 30 mins on a 90 core server')")
