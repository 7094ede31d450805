#!/usr/bin/env python
"""Generate tests/golden/reference_fixtures.npz from the reference's shipped .RData files.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py
The GPU box has no /root/reference, so the extracted arrays are committed.

Sources:
  lgssm/data/data100obs.RData          get_ar(1) dataset, theta=(0.5, sqrt(10))
  lgssm/data/lgssmlargem.RData         authors' Var(log Z) vs N (10^4 filters each)
  stochastic_kinetic/data/*.RData      SK dataset + authors' Var(log Z) and meeting times
  sv/sp500.RData, sv/data/parameter_mle.RData
  smcsamplers/data/data100obs.RData    get_gss dataset
Also stores the exact Kalman log-likelihood of the LGSSM dataset (known answer),
both with the full 0.5*log(2*pi) and with the reference's truncated 0.9189385
(src/mvnorm.cpp:121).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from unbiased_pimh_b200.rdata import read_rdata  # noqa: E402

REF = os.environ.get("CPIMH_REFERENCE", "/root/reference") + "/inst/reproduce/"


def kalman_loglik(y, alpha, sigma_y, half_log_2pi):
    """Exact log p(y_{1:T}) for x0~N(0,1), x_t = alpha^1 * x_{t-1} + N(0,1), y_t ~ N(x_t, sigma_y^2).

    get_ar(1): A = alpha^(1+|i-j|) = alpha (src/create_A_.cpp:10); rinit N(0,1)
    (R/model_get_ar.R:14,17); the first transition is applied before the first
    observation (R/CPF.R:44,60)."""
    m, P = 0.0, 1.0
    ll = 0.0
    for t in range(len(y)):
        m, P = alpha * m, alpha * alpha * P + 1.0
        S = P + sigma_y ** 2
        ll += -half_log_2pi - 0.5 * np.log(S) - 0.5 * (y[t] - m) ** 2 / S
        K = P / S
        m, P = m + K * (y[t] - m), (1 - K) * P
    return ll


def main():
    out = {}
    d = read_rdata(REF + "lgssm/data/data100obs.RData")
    out["lgssm_x"] = d["x"]
    out["lgssm_y"] = d["observations"]
    y = d["observations"][:, 0]
    out["lgssm_kalman_loglik"] = np.array(kalman_loglik(y, 0.5, np.sqrt(10.0), 0.5 * np.log(2 * np.pi)))
    out["lgssm_kalman_loglik_truncconst"] = np.array(kalman_loglik(y, 0.5, np.sqrt(10.0), 0.9189385))
    d = read_rdata(REF + "lgssm/data/lgssmlargem.RData")
    out["lgssm_varlogz_N"] = np.asarray(d["N_arr_unique"], dtype=np.float64)
    out["lgssm_varlogz"] = np.asarray(d["var_logz_res"], dtype=np.float64)
    out["lgssm_varlogz_nrep"] = np.asarray(d["n_var"], dtype=np.float64)

    d = read_rdata(REF + "stochastic_kinetic/data/data100obs.RData")
    out["sk_x"] = d["x"]
    out["sk_y"] = d["observations"]
    d = read_rdata(REF + "stochastic_kinetic/data/stochastic_kinetic_mt_varz.RData")
    out["sk_varlogz_N"] = np.asarray(d["N_arr"], dtype=np.float64)
    out["sk_varlogz"] = np.asarray(d["varlogz_arr"], dtype=np.float64)
    mt = np.asarray(d["mt_arr"])  # 5 x 10^4 meeting times
    out["sk_mt_mean"] = mt.mean(axis=1)
    out["sk_mt_nrep"] = np.array(mt.shape[1], dtype=np.float64)
    out["sk_mt_hist"] = np.stack([np.bincount(np.minimum(r.astype(int), 20), minlength=21) for r in mt])

    d = read_rdata(REF + "sv/sp500.RData")
    df = d["observationsDF"]
    ysv = np.asarray(df["y"] if isinstance(df, dict) else df[0], dtype=np.float64)
    out["sv_y"] = ysv.reshape(-1, 1)
    d = read_rdata(REF + "sv/data/parameter_mle.RData")
    out["sv_theta_mle"] = np.asarray(d["theta_mle"], dtype=np.float64)

    d = read_rdata(REF + "smcsamplers/data/data100obs.RData")
    out["gss_x"] = d["x"]
    out["gss_y"] = d["observations"]

    path = os.path.join(HERE, "reference_fixtures.npz")
    np.savez_compressed(path, **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.ravel(v)[0]))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
