// ccpf.cuh -- conditional particle filters, one CTA per filter:
//   CPF(N, model, theta, obs, ref_trajectory = x, with_as)                  R/CPF.R:14-78 (branches :22-24, :45-58)
//   CPF_coupled(N, model, theta, obs, ref1, ref2, CR_indexmatching, with_as) R/CPF_coupled.R:14-112
// The last particle of each system is pinned to its reference path; the other N-1 are resampled -- by sorted
// multinomial draws (one system, src/multinomial.cpp:13-32) or by index-matching coupled resampling (two systems,
// src/coupledresamplingindexmatching.cpp:8-57) -- and both systems are driven by the SAME Philox streams (common random
// numbers, R/CPF_coupled.R:24,29,57-59), so that particles that coincide stay bitwise identical and the two returned
// trajectories can meet exactly.  The sorted-uniform ladder, the coupling flags and the one-draw index matching used
// for ancestor sampling and the final path are specified in oracle/ccpf.c (header) and DESIGN.md "RNG".
//
// Layout: the particle history is the path store, as in the bootstrap filter's PATH_DENSE mode, once per system:
// hist_x[ws][sys][t][q][i] and hist_a[ws][sys][t][i].  Shared memory holds, per CTA, the two weight vectors (which become
// the residual CDFs), nu = min(w1, w2) (its CDF), the ladder of exponential-spacing prefix sums and the per-index
// coupled/residual ranks; the log-weights of a step reuse the nu / ladder arrays once all ancestors are known.
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace cpimh {

struct CcpfParams {
  int N, T, dy, Npad;
  long long B;                 // filters in this launch
  int nsys;                    // 1 or 2 (when `single` is null)
  int with_as;
  unsigned long long seed, first_filter;
  const unsigned long long* stream_ids;   // [B] explicit Philox stream id per filter, or null: first_filter + b
  const int* slot;             // [B] slot of refs / outputs used by filter b (chain driver: replicate), or null: b
  const int* single;           // [slot] != 0: run system 1 only (chain driver: the replicate's chains have met), or null
  const double* obs; const double* theta; const double* pre;
  int npre, integrator, substeps;
  const double* ref1; const double* ref2;   // [slot][T+1][D]
  double* hist_x;              // [ws][2][T+1][D][Npad]
  int* hist_a;                 // [ws][2][T][Npad]
  double* traj1; double* traj2;             // [slot][T+1][D]
  double* logz;                // [slot]   system 1's log Z (R/CPF.R:71)
  int* path_index;             // [slot][2] or null
  int* status;                 // [slot]
  int* anc_rec;                // [slot][2][T][N] or null
  double* exp1; double* exp2;  // [slot][T+1][D] or null: weighted average of all N paths (CPF_RB, R/rheeglynn_estimator_RB.R:50-56)
  const double* inj_init; const double* inj_trans;
  int n_init, n_trans, sk_M;
  // exact ancestors (DESIGN.md "Exact ancestors"): a step in which a uniform lies within margin_rel of a CDF knot (or of the
  // coupling probability alpha) is repeated inside the kernel in reference order (near[0] counts the repeats): alpha, mu = nu / alpha, the
  // residuals (w - nu) / (1 - alpha) and every cumsum in the reference's order (src/coupledresamplingindexmatching.cpp:13-24,
  // src/multinomial.cpp:17,40-41) by one thread
  int exact; int* near; double margin_rel;
  double* wsave;               // [ws][2][Npad] the step's normalised weights (L2-resident scratch, read only by a repeat)
};

struct CcpfSmem {
  SmemArr<double> C1, C2;      // [Npad] weights of system 1 / 2 -> residual CDFs
  SmemArr<double> V;           // [Npad + 32] nu -> its CDF ; log-weights of system 1 ; ancestor-sampling weights
  SmemArr<double> P;           // [Npad + 32] ladder prefix sums ; log-weights of system 2
  SmemArr<int> R;              // [Npad] rank among the coupled (>= 0) or the residual (~rank < 0) indices
  SmemArr<double> theta, pre, dscr;
  SmemArr<int> iscr;
};
__host__ __device__ inline size_t ccpf_smem_bytes(int Npad, int npre) {
  size_t b = (size_t)(4 * padded_len(Npad + 32)) * 8 + (size_t)(CPIMH_MAX_THETA + npre + (npre & 1) + 64) * 8 + (size_t)Npad * 4 + 36 * 4;
  return (b + 15) & ~(size_t)15;
}
__device__ __forceinline__ CcpfSmem ccpf_carve(int Npad, int npre) {
  CcpfSmem s;
  int b = 0;
  const int arr = padded_len(Npad + 32) * 8;   // all four in the padded layout (element i at pad16(i), common.cuh)
  s.C1.off = b; b += arr;
  s.C2.off = b; b += arr;
  s.V.off = b; b += arr;
  s.P.off = b; b += arr;
  s.theta.off = b; b += CPIMH_MAX_THETA * 8;
  s.pre.off = b; b += (npre + (npre & 1)) * 8;
  s.dscr.off = b; b += 64 * 8;
  s.R.off = b; b += Npad * 4;
  s.iscr.off = b;
  return s;
}

__device__ __forceinline__ int clampN(int a, int N) { return a >= N ? N - 1 : a; }

// in-place CDF of a padded array: parallel scan, or (exact) the reference's left-to-right running sum.  All threads call.
__device__ __forceinline__ void cc_cdf(double* A, int N, bool exact, double* dscr) {
  if (exact) { __syncthreads(); if (threadIdx.x == 0) sequential_cumsum_p(A, N); __syncthreads(); }
  else block_scan_chunked<false>(A, nullptr, N, dscr);
}
// sum of v over the CTA: parallel tree, or (exact) A[0] + A[1] + ... by one thread (Rcpp sugar sum)
__device__ __forceinline__ double cc_sum_min(const double* A, const double* Bv, int N, bool exact, double* dscr) {
  const int tid = threadIdx.x, NT = blockDim.x;
  if (exact) {
    __syncthreads();
    if (tid == 0) { double acc = 0.0; for (int k = 0; k < N; k++) acc += fmin(A[pad16(k)], Bv[pad16(k)]); dscr[0] = acc; }
    __syncthreads();
    const double r = dscr[0];
    __syncthreads();
    return r;
  }
  double al = 0.0;
  for (int k = tid; k < N; k += NT) al += fmin(A[pad16(k)], Bv[pad16(k)]);
  return block_sum(al, dscr);
}
// the two knots that decided index a against u +- the rounding margin; LE: rule `u < cdf[j-1]` (src/multinomial.cpp:25-27),
// else `while (csw < u)` (src/systematic.cpp:15)
template <bool LE>
__device__ __forceinline__ bool cc_near(const double* C, int N, int a, double u, double margin_rel) {
  const double mar = margin_rel * C[pad16(N - 1)];
  const double cur = a < N ? C[pad16(a)] : INFINITY, below = a > 0 ? C[pad16(a - 1)] : -INFINITY;
  return LE ? (cur - u <= mar || u - below < mar) : (cur - u < mar || u - below <= mar);
}

// indexmatching_cpp(w1, w2, ndraws = 1) with the uniforms (uc, ud): coupled with probability alpha = sum min(w1, w2),
// then one draw from mu, else one common-uniform draw from the two residuals.  Destroys A and Bv.  All threads call.
__device__ inline void im_single_draw(double* A, double* Bv, int N, double uc, double ud, double* dscr, int& a1, int& a2,
                                      bool exact, double margin_rel, bool& near) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const double al = cc_sum_min(A, Bv, N, exact, dscr);
  if (!exact && fabs(uc - al) <= margin_rel) near = true;
  if (uc < al) {
    for (int k = tid; k < N; k += NT) { const double nu = fmin(A[pad16(k)], Bv[pad16(k)]); A[pad16(k)] = exact ? nu / al : nu; }
    __syncthreads();
    cc_cdf(A, N, exact, dscr);
    const double u = ud * A[pad16(N - 1)];
    const int a = count_le_p(A, 0, N, u);
    if (!exact && cc_near<true>(A, N, a, u, margin_rel)) near = true;
    a1 = a2 = clampN(a, N);
  } else {
    const double den = 1.0 - al;
    for (int k = tid; k < N; k += NT) {
      const double nu = fmin(A[pad16(k)], Bv[pad16(k)]);
      A[pad16(k)] = exact ? (A[pad16(k)] - nu) / den : A[pad16(k)] - nu;
      Bv[pad16(k)] = exact ? (Bv[pad16(k)] - nu) / den : Bv[pad16(k)] - nu;
    }
    __syncthreads();
    if (exact) { cc_cdf(A, N, true, dscr); cc_cdf(Bv, N, true, dscr); }
    else block_scan_chunked<true>(A, Bv, N, dscr);
    const double u1 = ud * A[pad16(N - 1)], u2 = ud * Bv[pad16(N - 1)];
    const int b1 = count_le_p(A, 0, N, u1), b2 = count_le_p(Bv, 0, N, u2);
    if (!exact && (cc_near<true>(A, N, b1, u1, margin_rel) || cc_near<true>(Bv, N, b2, u2, margin_rel))) near = true;
    a1 = clampN(b1, N); a2 = clampN(b2, N);
  }
  __syncthreads();
}

template <class Model, int NSYS>
__device__ void run_ccpf(const CcpfParams& p, const CcpfSmem& sm, long long ws, const Stream& stream, long long inj, long long slot) {
  constexpr int D = Model::D;
  constexpr int nsys = NSYS;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int N = p.N, T = p.T, Npad = p.Npad, nd = N - 1;
  const size_t xstride = (size_t)D * Npad;
  double* hx[2];
  int* ha[2];
  const double* ref[2] = {p.ref1 + (size_t)slot * (T + 1) * D, p.ref2 ? p.ref2 + (size_t)slot * (T + 1) * D : nullptr};
#pragma unroll
  for (int s = 0; s < 2; s++) {
    hx[s] = p.hist_x + ((size_t)ws * 2 + s) * (T + 1) * xstride;
    ha[s] = p.hist_a + ((size_t)ws * 2 + s) * T * Npad;
  }
  const SmemArr<double> Wk[2] = {sm.C1, sm.C2};
  const SmemArr<double> Lk[2] = {sm.V, sm.P};
  ModelCtx c;
  c.theta = sm.theta; c.pre = sm.pre; c.stream = stream; c.N = N; c.T = T;
  c.inj_init = p.inj_init ? p.inj_init + (size_t)inj * p.n_init * N : nullptr;
  c.inj_trans = p.inj_trans ? p.inj_trans + (size_t)inj * T * p.n_trans * N : nullptr;
  c.n_trans = p.n_trans; c.sk_M = p.sk_M; c.integrator = p.integrator; c.substeps = p.substeps;
  int st = 0;
  if (tid == 0) sm.iscr[32] = 0;
  const bool exact = p.exact != 0;
  bool near = false;
  int repeats = 0;
  double* wsave[2] = {p.wsave + (size_t)ws * 2 * Npad, p.wsave + ((size_t)ws * 2 + 1) * Npad};

  // ---------------- initialisation (R/CPF.R:19-28, R/CPF_coupled.R:22-33): common init noise, last particle = ref[, 1]
  for (int i = tid; i < N; i += NT) {
    double x[D];
    Model::init(c, i, x);
#pragma unroll
    for (int s = 0; s < nsys; s++) {
#pragma unroll
      for (int q = 0; q < D; q++) hx[s][(size_t)q * Npad + i] = (i == N - 1) ? ref[s][q] : x[q];
      Wk[s][pad16(i)] = 1.0;
    }
  }
  double ssum[2] = {(double)N, (double)N};
  double logz_sum = 0.0;
  __syncthreads();

  for (int t = 1; t <= T; t++) {
    const bool ident = (t == 1) || isnan(p.obs[(size_t)(t - 2) * p.dy]);     // R/CPF.R:38-40, R/CPF_coupled.R:45-48
    const double* y = p.obs + (size_t)(t - 1) * p.dy;
    const double sc = Model::step_const(c, t);
    const double* xprev[2] = {hx[0] + (size_t)(t - 1) * xstride, hx[1] + (size_t)(t - 1) * xstride};
    double* xcur[2] = {hx[0] + (size_t)t * xstride, hx[1] + (size_t)t * xstride};
    int* arow[2] = {ha[0] + (size_t)(t - 1) * Npad, ha[1] + (size_t)(t - 1) * Npad};

    // normweights = w / sum(w)  (R/CPF.R:63; index matching compares the two systems' NORMALISED weights)
#pragma unroll
    for (int s = 0; s < nsys; s++)
      for (int k = tid; k < N; k += NT) Wk[s][pad16(k)] = Wk[s][pad16(k)] / ssum[s];
    __syncthreads();

    // the normalised weights, kept for a repeat of this step: the CDFs below overwrite them in place
#pragma unroll
    for (int s = 0; s < nsys; s++)
      for (int k = tid; k < N; k += NT) wsave[s][k] = Wk[s][pad16(k)];
    int aref[2] = {N - 1, N - 1};
    bool ex = exact;
    for (;;) {                                                               // second pass only after a near-knot draw (reference order)
    near = false;
    aref[0] = aref[1] = N - 1;
    // ---- ancestor of the reference particle: R/CPF.R:47-57, R/CPF_coupled.R:64-84
    if constexpr (Model::HAS_DTRANS) {
      if (p.with_as) {
#pragma unroll
        for (int s = 0; s < nsys; s++) {
          double next[D];
#pragma unroll
          for (int q = 0; q < D; q++) next[q] = ref[s][(size_t)t * D + q];
          double m = -INFINITY;
          for (int k = tid; k < N; k += NT) {
            double x[D];
#pragma unroll
            for (int q = 0; q < D; q++) x[q] = xprev[s][(size_t)q * Npad + k];
            const double lm = log(Wk[s][pad16(k)]) + Model::dtrans(c, next, x, t, sc);
            Lk[s][pad16(k)] = lm;
            m = fmax(m, lm);
          }
          m = block_max(m, sm.dscr);
          double sum = 0.0;
          for (int k = tid; k < N; k += NT) { const double w = exp(Lk[s][pad16(k)] - m); Lk[s][pad16(k)] = w; sum += w; }
          sum = block_sum(sum, sm.dscr);
          for (int k = tid; k < N; k += NT) Lk[s][pad16(k)] = Lk[s][pad16(k)] / sum;
        }
        __syncthreads();
        double ua, ub;
        philox_u2(stream, 0, t, CPIMH_P_AS, 0, ua, ub);
        if (nsys == 1) {
          cc_cdf(Lk[0], N, ex, sm.dscr);
          const int a = count_lt_p(Lk[0], N, ua);                              // systematic_resampling_n(w_as, 1, u)
          if (!ex && cc_near<false>(Lk[0], N, a, ua, p.margin_rel)) near = true;
          aref[0] = clampN(a, N);
          __syncthreads();
        } else {
          im_single_draw(Lk[0], Lk[1], N, ua, ub, sm.dscr, aref[0], aref[1], ex, p.margin_rel, near);  // CR_indexmatching(w_as1, w_as2, ntrials = 1)
        }
      }
    }

    // ---- ancestors of the N-1 free particles
    if (!ident) {
      for (int i = tid; i <= nd + 1; i += NT) sm.P[pad16(i)] = -rng_log(philox_u1(stream, i, t, CPIMH_P_RESAMPLE, 0));
      if (nsys == 1) {
        __syncthreads();
        cc_cdf(sm.C1, N, ex, sm.dscr);
        block_scan_chunked<false>(sm.P, nullptr, nd + 1, sm.dscr);
        const double scale = sm.C1[pad16(N - 1)] / sm.P[pad16(nd)];
        for (int k = tid; k < nd; k += NT) {
          const double u = scale * sm.P[pad16(k)];
          int a = count_le_p(sm.C1, 0, N, u);
          if (!ex && cc_near<true>(sm.C1, N, a, u, p.margin_rel)) near = true;
          if (a >= N) { a = N - 1; st |= 512; }
          arow[0][k] = a;
        }
      } else {
        const double al = cc_sum_min(sm.C1, sm.C2, N, ex, sm.dscr);       // coupledresamplingindexmatching.cpp:13-21
        {
          const double den = 1.0 - al;
          for (int k = tid; k < N; k += NT) {
            const double nu = fmin(sm.C1[pad16(k)], sm.C2[pad16(k)]);
            if (ex) {                                                     // mu = nu / alpha, R = (w - nu) / (1 - alpha)  (:22-24)
              sm.V[pad16(k)] = nu / al;
              sm.C1[pad16(k)] = al < 1.0 ? (sm.C1[pad16(k)] - nu) / den : 0.0;
              sm.C2[pad16(k)] = al < 1.0 ? (sm.C2[pad16(k)] - nu) / den : 0.0;
            } else { sm.V[pad16(k)] = nu; sm.C1[pad16(k)] -= nu; sm.C2[pad16(k)] -= nu; }
          }
        }
        // coupled flags (:26-33) and each index's rank inside its group: thread `tid` owns a contiguous chunk of indices
        const int chunk = (nd + NT - 1) / NT, i0 = min(nd, tid * chunk), i1 = min(nd, i0 + chunk);
        int cnt = 0;
        for (int i = i0; i < i1; i++) {
          const double uc = philox_u1(stream, i, t, CPIMH_P_COUPLE, 0);
          const int f = uc < al;
          if (!ex && fabs(uc - al) <= p.margin_rel) near = true;
          sm.R[i] = f; cnt += f;
        }
        int nc;
        int jc = block_excl_scan_int(cnt, sm.iscr, &nc);
        int jr = i0 - jc;
        for (int i = i0; i < i1; i++) { if (sm.R[i]) sm.R[i] = jc++; else sm.R[i] = ~(jr++); }
        if (ex) { cc_cdf(sm.C1, N, true, sm.dscr); cc_cdf(sm.C2, N, true, sm.dscr); cc_cdf(sm.V, N, true, sm.dscr); }
        else { block_scan_chunked<true>(sm.C1, sm.C2, N, sm.dscr); block_scan_chunked<false>(sm.V, nullptr, N, sm.dscr); }
        block_scan_chunked<false>(sm.P, nullptr, nd + 2, sm.dscr);
        const double Pnc = sm.P[pad16(nc)], den = sm.P[pad16(nd + 1)] - Pnc;
        const double totV = sm.V[pad16(N - 1)], tot1 = sm.C1[pad16(N - 1)], tot2 = sm.C2[pad16(N - 1)];
        for (int k = tid; k < nd; k += NT) {
          const int r = sm.R[k];
          int a1, a2;
          if (r >= 0) {                                                      // :36-38 common draw from mu = nu / alpha
            const double u = (sm.P[pad16(r)] / Pnc) * totV;
            const int a = count_le_p(sm.V, 0, N, u);
            if (!ex && cc_near<true>(sm.V, N, a, u, p.margin_rel)) near = true;
            a1 = a2 = clampN(a, N);
          } else {                                                           // :40-42 residuals with a common uniform
            const double u = (sm.P[pad16(nc + 1 + (~r))] - Pnc) / den;
            const int b1 = count_le_p(sm.C1, 0, N, u * tot1), b2 = count_le_p(sm.C2, 0, N, u * tot2);
            if (!ex && (cc_near<true>(sm.C1, N, b1, u * tot1, p.margin_rel) || cc_near<true>(sm.C2, N, b2, u * tot2, p.margin_rel))) near = true;
            a1 = clampN(b1, N); a2 = clampN(b2, N);
          }
          arow[0][k] = a1; arow[1][k] = a2;
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < nsys; s++)
        for (int k = tid; k < nd; k += NT) arow[s][k] = k;
    }
    if (!ex && __syncthreads_or(near ? 1 : 0)) {
      ex = true; repeats++;
#pragma unroll
      for (int s = 0; s < nsys; s++)
        for (int k = tid; k < N; k += NT) Wk[s][pad16(k)] = wsave[s][k];
      __syncthreads();
      continue;
    }
    break;
    }
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < nsys; s++) arow[s][N - 1] = aref[s];
    }
    __syncthreads();                                                         // ancestors complete: CDFs, ladder and ranks are dead

    // ---- gather -> common-noise transition -> pin the reference -> weigh  (R/CPF_coupled.R:50-62, 87-88)
    double m[2] = {-INFINITY, -INFINITY};
    for (int k = tid; k < N; k += NT) {
      uint2 ush = make_uint2(0u, 0u);
      if (Model::USES_SHARED_U && !c.inj_trans) { double ua; philox_u1raw(stream, k, t, CPIMH_P_RESAMPLE, 0, ua, ush); }
#pragma unroll
      for (int s = 0; s < nsys; s++) {
        double x[D];
        if (k == N - 1) {
#pragma unroll
          for (int q = 0; q < D; q++) x[q] = ref[s][(size_t)t * D + q];
        } else {
          const int a = arow[s][k];
#pragma unroll
          for (int q = 0; q < D; q++) x[q] = xprev[s][(size_t)q * Npad + a];
          st |= Model::transition(c, t, k, x, sc, ush);
        }
#pragma unroll
        for (int q = 0; q < D; q++) xcur[s][(size_t)q * Npad + k] = x[q];
        const double lw = Model::logw(c, x, y);
        Lk[s][pad16(k)] = lw;
        m[s] = fmax(m[s], lw);
      }
    }
    __syncthreads();
    // ---- weights and log Z increment (R/CPF.R:61-63,66)
#pragma unroll
    for (int s = 0; s < nsys; s++) {
      const double ms = block_max(m[s], sm.dscr);
      double sum = 0.0;
      for (int k = tid; k < N; k += NT) { const double w = exp(Lk[s][pad16(k)] - ms); Wk[s][pad16(k)] = w; sum += w; }
      ssum[s] = block_sum(sum, sm.dscr);
      if (s == 0) logz_sum += log(ssum[0] / (double)N) + ms;
    }
    __syncthreads();
  }

  // ---------------- path selection: R/CPF.R:69 (one system) / R/CPF_coupled.R:101-106 (one index-matching draw)
  if (st) atomicOr(&sm.iscr[32], st);
#pragma unroll
  for (int s = 0; s < nsys; s++)
    for (int k = tid; k < N; k += NT) Wk[s][pad16(k)] = Wk[s][pad16(k)] / ssum[s];
  __syncthreads();
  if (p.exp1) {
    // estimate = sum_k normweights[k] * Tree$get_path(k) (R/rheeglynn_estimator_RB.R:50-56, test_function = identity): all N
    // leaves walk to the root in lock step, one block reduction per generation and component; R holds the walkers
#pragma unroll
    for (int s = 0; s < nsys; s++) {
      double* ex = (s == 0 ? p.exp1 : p.exp2) + (size_t)slot * (T + 1) * D;
      for (int n = tid; n < N; n += NT) sm.R[n] = n;
      for (int step = T; step >= 0; step--) {
#pragma unroll
        for (int q = 0; q < D; q++) {
          double acc = 0.0;
          for (int n = tid; n < N; n += NT) acc += hx[s][(size_t)step * xstride + (size_t)q * Npad + sm.R[n]] * Wk[s][pad16(n)];
          acc = block_sum(acc, sm.dscr);
          if (tid == 0) ex[(size_t)step * D + q] = acc;
        }
        if (step > 0) for (int n = tid; n < N; n += NT) sm.R[n] = ha[s][(size_t)(step - 1) * Npad + sm.R[n]];
      }
    }
    __syncthreads();
  }
  double ua, ub;
  philox_u2(stream, 0, T + 1, CPIMH_P_PATH, 0, ua, ub);
  int kidx[2] = {0, 0};
#pragma unroll
  for (int s = 0; s < nsys; s++)
    for (int k = tid; k < N; k += NT) wsave[s][k] = Wk[s][pad16(k)];
  for (bool ex = exact;;) {
    near = false;
    if (nsys == 1) {
      cc_cdf(sm.C1, N, ex, sm.dscr);
      const int a = count_lt_p(sm.C1, N, ua);
      if (!ex && cc_near<false>(sm.C1, N, a, ua, p.margin_rel)) near = true;
      kidx[0] = clampN(a, N);
    } else {
      im_single_draw(sm.C1, sm.C2, N, ua, ub, sm.dscr, kidx[0], kidx[1], ex, p.margin_rel, near);
    }
    if (!ex && __syncthreads_or(near ? 1 : 0)) {
      ex = true; repeats++;
#pragma unroll
      for (int s = 0; s < nsys; s++)
        for (int k = tid; k < N; k += NT) Wk[s][pad16(k)] = wsave[s][k];
      __syncthreads();
      continue;
    }
    break;
  }
  if (tid == 0 && repeats && p.near) atomicAdd(p.near, repeats);
  const int flags = sm.iscr[32];
  if (tid == 0) {
    p.logz[slot] = logz_sum;
    if (p.path_index) { p.path_index[2 * slot] = kidx[0]; p.path_index[2 * slot + 1] = nsys == 2 ? kidx[1] : -1; }
    p.status[slot] = flags & 255;
  }
  if (tid < 32) {                                      // Tree$get_path: one warp walks the ancestor rows, lane q carries component q
#pragma unroll
    for (int s = 0; s < nsys; s++) {
      double* tr = (s == 0 ? p.traj1 : p.traj2) + (size_t)slot * (T + 1) * D;
      int j = kidx[s];
      for (int step = T; step >= 0; step--) {
        for (int q = tid; q < D; q += 32) tr[(size_t)step * D + q] = hx[s][(size_t)step * xstride + (size_t)q * Npad + j];
        if (step > 0) j = ha[s][(size_t)(step - 1) * Npad + j];
      }
    }
  }
  if (p.anc_rec)
#pragma unroll
    for (int s = 0; s < nsys; s++)
      for (int t = 0; t < T; t++)
        for (int k = tid; k < N; k += NT) p.anc_rec[(((size_t)slot * 2 + s) * T + t) * N + k] = ha[s][(size_t)t * Npad + k];
  __syncthreads();
}

template <class Model>
__global__ void __launch_bounds__(512, 1) ccpf_kernel(CcpfParams p) {
  CcpfSmem sm = ccpf_carve(p.Npad, p.npre);
  for (int i = threadIdx.x; i < CPIMH_MAX_THETA; i += blockDim.x) sm.theta[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.npre; i += blockDim.x) sm.pre[i] = p.pre[i];
  __syncthreads();
  for (long long b = blockIdx.x; b < p.B; b += gridDim.x) {
    const long long slot = p.slot ? p.slot[b] : b;
    const int nsys = p.single ? (p.single[slot] ? 1 : 2) : p.nsys;
    const Stream s = make_stream(p.seed, p.stream_ids ? p.stream_ids[b] : p.first_filter + (unsigned long long)b);
    if (nsys == 1) run_ccpf<Model, 1>(p, sm, blockIdx.x, s, b, slot);
    else run_ccpf<Model, 2>(p, sm, blockIdx.x, s, b, slot);
  }
}

// ------------------------------------------------------------------ coupled_ccpf_chains state machine
struct CcRound {
  int PD, k, K, max_it, it;
  unsigned first_replicate;
  const int* active; int nA; int* next_active; int* next_count;
  unsigned long long* stream_ids;
  const double* new1; const double* new2; const double* new_logz; const int* new_status;
  double* st1; double* st2; double* lz1; double* lz2;
  int* single; int* iter; int* tau; int* status;
  double* accH; double* accD; double* hbar; double* bc;
  unsigned long long* nfilters;
  // Rao-Blackwellised twin (null when off): exp paths of the new filters / of the current states, accumulators, output
  const double* newe1; const double* newe2; double* e1; double* e2; double* accHe; double* accDe; double* hbar_rb;
  // Rhee-Glynn estimator with geometric truncation (R/rheeglynn_estimator.R:24-31, 58-59): per-replicate truncation variable (null:
  // none) and the parameter lambda of its law; difference n is divided by the survival probability (1 - lambda)^n
  const int* trunc; double rg_lambda;
};
// one step of the lag-one recursion of replicate `rep` after its iteration-`it` filter(s) (R/debiasing_algorithm.R:50-106, 296-312);
// all threads of the CTA call; returns != 0 when the replicate is finished
__device__ inline int cc_accept(const CcRound& q, int rep, int it) {
  const int PD = q.PD;
  double* x1 = q.st1 + (size_t)rep * PD;
  double* x2 = q.st2 + (size_t)rep * PD;
  const double* n1 = q.new1 + (size_t)rep * PD;
  const double* n2 = q.new2 + (size_t)rep * PD;
  double* accH = q.accH + (size_t)rep * PD;
  double* accD = q.accD + (size_t)rep * PD;
  const int single = q.single[rep];
  int tau = q.tau[rep];
  int status = q.new_status[rep];
  if (status == 0 && isnan(q.new_logz[rep])) status = CPIMH_ERR_INVALID;
  const int t = it - 1;
  const bool mcmc = (it >= q.k && it <= q.K);
  double coef = (double)min(t - q.k + 1, q.K - q.k + 1);
  if (q.rg_lambda > 0.0) coef = coef / pow(1.0 - q.rg_lambda, (double)it);             // / survival_probabilities[iteration]
  int same = 1;
  for (int j = threadIdx.x; j < PD; j += blockDim.x) {
    const double xa = n1[j];
    // iteration 1 moves chain 1 only; after the meeting both chains take the single kernel's output (:68-73)
    const double xb = (it == 1) ? x2[j] : (single ? xa : n2[j]);
    x1[j] = xa; x2[j] = xb;
    if (xa != xb) same = 0;
  }
  same = __syncthreads_and(same);
  int finished = 0;
  if (status) finished = 1;
  if (it >= 2 && !single && same && tau < 0) tau = it;                         // :82-86 all(chain_state1 == chain_state2)
  if (tau > 0 && it >= max(tau, q.K)) finished = 1;                            // :104-106
  if (!finished && q.trunc && it >= q.trunc[rep]) finished = 1;                          // the truncation variable ends the sum (:45)
  if (!finished && q.max_it >= 0 && it >= q.max_it) { finished = 1; status = CPIMH_ERR_NOT_MET; }   // R: finished = FALSE
  // H_bar (debiasing_algorithm.R:296-312): MCMC sum over k..K plus coef_t (X_{t+1} - Y_t) for k <= t <= tau - 1, skipped
  // altogether when tau <= k + 1 (:303).  For the states the skipped term is zero; for the weighted-average paths it is not.
  const bool bias = (t >= q.k) && !(tau > 0 && tau <= q.k + 1);
  for (int j = threadIdx.x; j < PD; j += blockDim.x) {
    const double xa = x1[j], xb = x2[j];
    if (mcmc) accH[j] += xa;
    if (bias) accD[j] = accD[j] + coef * (xa - xb);
    if (q.e1) {                                                                // the same recursion over the filters' weighted-average paths
      const size_t o = (size_t)rep * PD + j;
      const double ea = q.newe1[o];
      const double eb = (it == 1) ? q.e2[o] : (single ? ea : q.newe2[o]);
      q.e1[o] = ea; q.e2[o] = eb;
      if (mcmc) q.accHe[o] += ea;
      if (bias) q.accDe[o] = q.accDe[o] + coef * (ea - eb);
    }
  }
  if (finished) {
    const double denom = (double)(q.K - q.k + 1);
    for (int j = threadIdx.x; j < PD; j += blockDim.x) {
      q.hbar[(size_t)rep * PD + j] = (accH[j] + accD[j]) / denom;
      q.bc[(size_t)rep * PD + j] = accD[j] / denom;
      if (q.e1) q.hbar_rb[(size_t)rep * PD + j] = (q.accHe[(size_t)rep * PD + j] + q.accDe[(size_t)rep * PD + j]) / denom;
    }
  }
  if (threadIdx.x == 0) {
    if (it == 1 || single) { q.lz1[rep] = q.new_logz[rep]; if (it > 1) q.lz2[rep] = q.new_logz[rep]; }   // coupled kernel leaves log_pdf unchanged
    q.single[rep] = (tau > 0) ? 1 : 0;
    q.iter[rep] = it; q.tau[rep] = tau;
    if (status) q.status[rep] = status;
    if (finished) atomicAdd(q.nfilters, (unsigned long long)(2 + it));
    else if (q.next_active) q.next_active[atomicAdd(q.next_count, 1)] = rep;
  }
  return finished;
}

// coupled_ccpf_chains as ONE persistent kernel: a CTA takes a replicate from the ticket counter and runs its whole chain --
// conditional filter(s) on the current states, accept / meet / H_bar recursion -- until iteration max(tau, K); no host round
// trip per iteration (round 1: a cudaMemcpy + stream synchronise after every iteration, 0.87 scaling efficiency at 8 GPUs).
// Unlike PIMH, a conditional filter depends on the chain state, so a replicate's iterations are inherently sequential.
template <class Model>
__global__ void __launch_bounds__(512, 1) ccpf_chains_kernel(CcpfParams p, CcRound q, long long R, int* work_counter) {
  CcpfSmem sm = ccpf_carve(p.Npad, p.npre);
  for (int i = threadIdx.x; i < CPIMH_MAX_THETA; i += blockDim.x) sm.theta[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.npre; i += blockDim.x) sm.pre[i] = p.pre[i];
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) sm.iscr[33] = atomicAdd(work_counter, 1);
    __syncthreads();
    const long long r = sm.iscr[33];
    __syncthreads();
    if (r >= R) break;
    if (q.status[r] != 0) continue;                          // an initial bootstrap filter failed
    if (q.trunc && q.trunc[r] <= 0) {                        // truncation = 0: the estimate is X_0 alone (R/rheeglynn_estimator.R:37-39)
      const double denom = (double)(q.K - q.k + 1);
      for (int j = threadIdx.x; j < q.PD; j += blockDim.x) {
        q.hbar[(size_t)r * q.PD + j] = q.accH[(size_t)r * q.PD + j] / denom; q.bc[(size_t)r * q.PD + j] = 0.0;
        if (q.e1) q.hbar_rb[(size_t)r * q.PD + j] = q.accHe[(size_t)r * q.PD + j] / denom;
      }
      if (threadIdx.x == 0) atomicAdd(q.nfilters, 2ull);
      continue;
    }
    for (int it = 1;; it++) {
      const Stream s = make_stream(p.seed, (unsigned long long)(q.first_replicate + (unsigned)r) | ((unsigned long long)it << 32));
      if (q.single[r]) run_ccpf<Model, 1>(p, sm, blockIdx.x, s, r, r);
      else run_ccpf<Model, 2>(p, sm, blockIdx.x, s, r, r);
      __syncthreads();
      const int fin = cc_accept(q, (int)r, it);
      __syncthreads();
      if (fin) break;
    }
  }
}

struct CcpfLaunch { int grid, threads; size_t smem; };
template <class Model>
int launch_ccpf(const CcpfParams& p, const CcpfLaunch& L, cudaStream_t stream) {
  if (p.with_as && !Model::HAS_DTRANS) { set_error("ancestor sampling needs model$dtransition: defined for get_gss and get_ar only"); return CPIMH_ERR_INVALID; }
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(ccpf_kernel<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  ccpf_kernel<Model><<<L.grid, L.threads, L.smem, stream>>>(p);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}
template <class Model>
int launch_ccpf_chains(const CcpfParams& p, const CcRound& q, long long R, int* work_counter, const CcpfLaunch& L, cudaStream_t stream) {
  if (p.with_as && !Model::HAS_DTRANS) { set_error("ancestor sampling needs model$dtransition: defined for get_gss and get_ar only"); return CPIMH_ERR_INVALID; }
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(ccpf_chains_kernel<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
  ccpf_chains_kernel<Model><<<L.grid, L.threads, L.smem, stream>>>(p, q, R, work_counter);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}
template <class Model>
int occupancy_ccpf(int threads, size_t smem, int* blocks_per_sm) {
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(ccpf_kernel<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CPIMH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ccpf_kernel<Model>, threads, smem));
  return CPIMH_OK;
}

}  // namespace cpimh
