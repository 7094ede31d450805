// models.cuh -- device-side rinit / rtransition / dmeasurement of the five shipped models, one particle
// per thread, state in registers.  Follows the specification the CPU oracle restates:
//   get_gss                 reference R/model_get_gss.R:13-46
//   get_ar(d)               R/model_get_ar.R:13-39, src/create_A_.cpp:6-13, src/mvnorm.cpp:113-132
//   get_levy_sv             R/model_levy_sv.R:17-66, src/SVLevy_functions.cpp:40-81
//   get_stochastic_kinetic  R/model_stochastic_kinetic.R:16-157 (exact Gillespie, 8 reactions)
//   get_lorenz              R/model_get_lorenz.R:18-50, src/lorenz96.cpp:13-32,65-93
// Noise comes from the filter's Philox streams (DESIGN.md "RNG") or from injected arrays.
#pragma once
#include "common.cuh"

namespace cpimh {

struct ModelCtx {
  SmemArr<double> theta;    // shared-memory copy of cfg.theta (dynamic window offset, see common.cuh SmemArr)
  SmemArr<double> pre;      // shared-memory model precompute (AR: A col-major then cst; LORENZ: cst)
  Stream stream;
  int N;                    // particles
  int T;
  const double* inj_init;   // this filter's injected init noise [n_init][N] or nullptr
  const double* inj_trans;  // this filter's injected transition noise [T][n_trans][N] or nullptr
  int n_trans;
  int sk_M;                 // SK injected events per particle-step
  int integrator, substeps; // LORENZ
};

__device__ __forceinline__ double dnorm_log(double x, double mu, double sigma) {   // R nmath/dnorm.c, give_log
  double z = (x - mu) / sigma;
  return -(CPIMH_M_LN_SQRT_2PI + 0.5 * z * z + log(sigma));
}
__device__ __forceinline__ const double* trans_row(const ModelCtx& c, int t, int row) {
  return c.inj_trans + ((size_t)(t - 1) * c.n_trans + row) * c.N;
}

// ------------------------------------------------------------------ get_gss (d = 1)
struct GssModel {
  static constexpr int D = 1;
  static constexpr bool USES_SHARED_U = false;
  static constexpr bool MULTI = false;
  static constexpr bool JUMP_LIST = false;
  static constexpr bool HAS_DTRANS = true;
  static constexpr bool SPLIT_W = false;
  // dtransition(next_x, xparticles, theta, time): R/model_get_gss.R:30-33
  __device__ static double dtrans(const ModelCtx&, const double* next, const double* x, int, double drift) {
    const double xi = x[0];
    const double means = 0.5 * xi + 25.0 * xi / (1.0 + xi * xi) + drift;
    return dnorm_log(next[0], means, 1.0);
  }
  __device__ static double step_const(const ModelCtx&, int t) { return 8.0 * cos(1.2 * (double)(t - 1)); }
  __device__ static void init(const ModelCtx& c, int i, double* x) {
    double z;
    if (c.inj_init) z = c.inj_init[i];
    else { double ua, ub; philox_u2(c.stream, i, 0, CPIMH_P_INIT, 0, ua, ub); z = normal1(ua, ub); }
    x[0] = sqrt(2.0) * z;
  }
  __device__ static int transition(const ModelCtx& c, int t, int i, double* x, double drift, uint2 /*zw*/) {
    double z;
    if (c.inj_trans) z = trans_row(c, t, 0)[i];
    else { double ua, ub; philox_u2(c.stream, i, t, CPIMH_P_TRANS, 0, ua, ub); z = normal1(ua, ub); }
    double xi = x[0];
    double means = 0.5 * xi + 25.0 * xi / (1.0 + xi * xi) + drift;
    x[0] = means + z;
    return 0;
  }
  __device__ static double logw(const ModelCtx&, const double* x, const double* y) {
    // dnorm(y, x^2 / 20, sd = sqrt(10), log = TRUE) (R/model_get_gss.R:36) with log(sd) folded into a constant
    const double z = (y[0] - (x[0] * x[0]) / 20.0) * CPIMH_INV_SQRT_10;
    return -(CPIMH_M_LN_SQRT_2PI + 0.5 * z * z + CPIMH_LOG_SQRT_10);
  }
};

// ------------------------------------------------------------------ get_ar(d)
template <int D_>
struct ArModel {
  static constexpr int D = D_;
  static constexpr bool USES_SHARED_U = false;   // pre = A (column-major, D*D) then cst
  static constexpr bool MULTI = false;
  static constexpr bool JUMP_LIST = false;
  static constexpr bool HAS_DTRANS = true;
  static constexpr bool SPLIT_W = false;
  // y = A x with A[i][j] = alpha^(1 + |i - j|) (src/create_A_.cpp:10, R/model_get_ar.R:13).  For D >= 3 the product uses the
  // structure of A instead of its D*D entries: with L_i = sum_{j<i} alpha^(i-j) x_j = alpha (x_{i-1} + L_{i-1}) and
  // U_i = sum_{j>i} alpha^(j-i) x_j = alpha (x_{i+1} + U_{i+1}),  y_i = alpha (x_i + L_i + U_i): 7D - 4 operations and no table
  // loads instead of D*D multiply-adds and D*D shared-memory loads (D = 10: 66 against 200 instructions; rounding differs from
  // the dense product by a few ulp, far inside the 1e-10 of the floating-point parity tests).
  __device__ static void apply_A(const ModelCtx& c, const double* x, double* y) {
    if (D < 3) {
#pragma unroll
      for (int r = 0; r < D; r++) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < D; l++) s += c.pre[r + l * D] * x[l];
        y[r] = s;
      }
    } else {
      const double a = c.pre[0];                  // A[0][0] = alpha
      double lo[D];
      double L = 0.0, U = 0.0;
      lo[0] = 0.0;
#pragma unroll
      for (int i = 1; i < D; i++) { L = a * (x[i - 1] + L); lo[i] = L; }
      y[D - 1] = a * (x[D - 1] + lo[D - 1]);
#pragma unroll
      for (int i = D - 2; i >= 0; i--) { U = a * (x[i + 1] + U); y[i] = a * ((x[i] + lo[i]) + U); }
    }
  }
  // dtransition = dmvnorm_transpose_cholesky(A x, next_x, I): R/model_get_ar.R:28-30, src/mvnorm.cpp:113-132
  __device__ static double dtrans(const ModelCtx& c, const double* next, const double* x, int, double) {
    double m[D];
    apply_A(c, x, m);
    double sq = 0.0;
#pragma unroll
    for (int r = 0; r < D; r++) { const double e = m[r] - next[r]; sq += e * e; }
    return -0.5 * sq + (-0.0 - (double)D * 0.9189385);
  }
  __device__ static double step_const(const ModelCtx&, int) { return 0.0; }
  __device__ static void normals(const ModelCtx& c, int i, int t, int purpose, double* z) {
#pragma unroll
    for (int j = 0; 2 * j < D; j++) {
      double ua, ub, z0, z1;
      philox_u2(c.stream, i, t, purpose, j, ua, ub);
      normal2(ua, ub, z0, z1);
      z[2 * j] = z0;
      if (2 * j + 1 < D) z[2 * j + 1] = z1;
    }
  }
  __device__ static void init(const ModelCtx& c, int i, double* x) {
    if (c.inj_init) {
#pragma unroll
      for (int k = 0; k < D; k++) x[k] = c.inj_init[(size_t)k * c.N + i];
    } else normals(c, i, 0, CPIMH_P_INIT, x);
  }
  __device__ static int transition(const ModelCtx& c, int t, int i, double* x, double, uint2 /*zw*/) {
    double z[D], y[D];
    if (c.inj_trans) {
#pragma unroll
      for (int k = 0; k < D; k++) z[k] = trans_row(c, t, k)[i];
    } else normals(c, i, t, CPIMH_P_TRANS, z);
    apply_A(c, x, y);
#pragma unroll
    for (int r = 0; r < D; r++) x[r] = y[r] + z[r];
    return 0;
  }
  __device__ static double logw(const ModelCtx& c, const double* x, const double* y) {
    const double isy = c.pre[D * D + 1];                  // 1 / theta[2]: the triangular solve of src/mvnorm.cpp:127 with L = sy I
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) { double r = (x[k] - y[k]) * isy; sq += r * r; }
    return -0.5 * sq + c.pre[D * D];
  }
};

// ------------------------------------------------------------------ get_levy_sv (d = 2; x = (v, z))
// Host precompute (api.cu, libm like the oracle): pre[0] = exp(-lambda), pre[1] = 1/lambda, pre[2] = omega^2/xi
// (scale of the exponential jumps), pre[3] = L, pre[4..4+L) = F_0..F_{L-1}: the CDF of rpois(lambda xi^2/omega^2)
// built by the recurrence p_k = p_{k-1} rate / k, F_k = F_{k-1} + p_k until it stops increasing; three guard knots; then
// pre[4+L+3 ..) = 1/p_0 .. 1/p_{L-1}.
// RNG layout (DESIGN.md "RNG"): the Philox block whose first half is the resampling draw of slot i at time t carries, in its
// second half, TWO 32-bit uniforms for the particle's transition: u_z -> the Poisson count k by inversion AND, through the
// position of u_z inside its cell, c'_1 = (u_z - F_{k-1}) / p_k -- uniform on (0,1) and independent of k -- the time of the first
// jump; u_w -> the size of the first jump.  69 % of the particles have no jump and 25 % exactly one: they need no further Philox
// call.  Jumps j >= 2 use block (i, t, TRANS, j) as before.  (R's own unif_rand() has 32-bit resolution.)
#define CPIMH_MULTI_MAX_E 16
__device__ __forceinline__ int poisson_tab(const double* pre, double u) {
  const int L = (int)pre[3];
  const double* F = pre + 4;
  // the first three knots without a dependent loop (the table is padded with three entries > 1 behind F_{L-1}); for the
  // rates of interest they hold > 99 % of the mass
  int k = (u > F[0] ? 1 : 0) + (u > F[1] ? 1 : 0) + (u > F[2] ? 1 : 0);
  if (k == 3) { while (k < L && u > F[k]) k++; }
  return k >= L ? 1000 : k;          // beyond the saturated CDF the sequential search only stops at its cap of 1000
}
struct LevySvModel {
  static constexpr int D = 2;
  static constexpr bool USES_SHARED_U = true;  // consumes the second half of the particle's resampling block (see above)
  static constexpr bool MULTI = false;
  static constexpr bool JUMP_LIST = true;      // pf_fused.cuh queues the particles that jump and walks the queue with full warps
  static constexpr bool HAS_DTRANS = false;
  static constexpr bool SPLIT_W = true;       // the weight factorises as exp(a) * f without a log (see weight_parts)
  __device__ static __forceinline__ int token(const ModelCtx& c, uint2 zw) { return poisson_tab(c.pre, u32_to_unit(zw.x)); }
  __device__ static double step_const(const ModelCtx& c, int) { return c.pre[0]; }
  // first jump of a particle with k >= 1 jumps, from the second half of its resampling block
  __device__ static __forceinline__ void jump1(const ModelCtx& c, uint2 zw, int k, double& wa, double& e) {
    const int L = (int)c.pre[3];
    const double uz = u32_to_unit(zw.x);
    const double fprev = k >= 2 ? c.pre[4 + k - 1] : (k == 1 ? c.pre[4] : 0.0);            // F_{k-1}
    const double cp = (uz - fprev) * c.pre[4 + L + 3 + (k < L ? k : L - 1)];            // c' = t - c ~ U(0,1)
    e = -rng_log(u32_to_unit(zw.y)) * c.pre[2];
    wa = exp(-c.theta[4] * cp) * e;
  }
  // jump j >= 2 of the compound Poisson process (src/SVLevy_functions.cpp:62-67): c' = t - c ~ U(0,1), e ~ Exp(rate xi/w2)
  __device__ static __forceinline__ void jump(const ModelCtx& c, int i, int t, int j, double& wa, double& e) {
    double ua, ub;
    philox_u2(c.stream, i, t, CPIMH_P_TRANS, j, ua, ub);
    e = -rng_log(ub) * c.pre[2];
    wa = exp(-c.theta[4] * ua) * e;
  }
  __device__ static void jump_sums(const ModelCtx& c, int i, int t, uint2 zw, double& swe, double& se) {
    const int k = poisson_tab(c.pre, u32_to_unit(zw.x));
    double a = 0.0, b = 0.0;
    if (k >= 1) jump1(c, zw, k, a, b);
    for (int j = 2; j <= k; j++) {
      double wa, e;
      jump(c, i, t, j, wa, e);
      a += wa;
      b += e;
    }
    swe = a; se = b;
  }
  __device__ static double gamma_mt(const ModelCtx& c, int i, double shape, double scale) {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
    for (int m = 0; m < 64; m++) {
      double ua, ub, z, z1, u2, ub2;
      philox_u2(c.stream, i, 0, CPIMH_P_INIT, 2 * m, ua, ub);
      normal2(ua, ub, z, z1);
      double v = 1.0 + cc * z;
      if (v <= 0) continue;
      v = v * v * v;
      philox_u2(c.stream, i, 0, CPIMH_P_INIT, 2 * m + 1, u2, ub2);
      if (log(u2) < 0.5 * z * z + d - d * v + d * log(v)) {
        double g = d * v;
        if (shape < 1.0) g *= pow(ub2, 1.0 / shape);
        return g * scale;
      }
    }
    return shape * scale;
  }
  __device__ static void init(const ModelCtx& c, int i, double* x) {
    const double xi = c.theta[2], w2 = c.theta[3];
    double z0, swe, se;
    if (c.inj_init) { z0 = c.inj_init[i]; swe = c.inj_init[(size_t)c.N + i]; se = c.inj_init[(size_t)2 * c.N + i]; }
    else {
      z0 = gamma_mt(c, i, xi * xi / w2, w2 / xi);
      double ua; uint2 zw;
      philox_u1raw(c.stream, i, 0, CPIMH_P_RESAMPLE, 0, ua, zw);
      jump_sums(c, i, 0, zw, swe, se);
    }
    double z1 = c.pre[0] * z0 + swe;
    x[0] = c.pre[1] * (z0 - z1 + se);
    x[1] = z1;
  }
  __device__ static __forceinline__ void finish(const ModelCtx& c, double* x, double swe, double se) {
    double zold = x[1];
    double znew = c.pre[0] * zold + swe;                 // src/SVLevy_functions.cpp:66
    x[0] = c.pre[1] * (zold - znew + se);                // :67
    x[1] = znew;
  }
  __device__ static int transition(const ModelCtx& c, int t, int i, double* x, double, uint2 zw) {
    double swe, se;
    if (c.inj_trans) { swe = trans_row(c, t, 0)[i]; se = trans_row(c, t, 1)[i]; }
    else jump_sums(c, i, t, zw, swe, se);
    finish(c, x, swe, se);
    return 0;
  }
  __device__ static double logw(const ModelCtx& c, const double* x, const double* y) {
    // dnorm(y, mu + beta v, sd = sqrt(v), log = TRUE) = -(M_LN_SQRT_2PI + (y - m)^2 / (2 v) + log(v) / 2)
    const double r = y[0] - (c.theta[0] + c.theta[1] * x[0]);
    return -(CPIMH_M_LN_SQRT_2PI + 0.5 * (r * r) / x[0] + 0.5 * log(x[0]));
  }
  // the same density as weight = exp(a) * f * exp(W_CONST): a = -(y - m)^2 / (2 v), f = v^(-1/2).  The filter only needs
  // w / sum(w) and log(mean(w)), so the per-particle log(v) (and the exp that undoes it) is replaced by one rsqrt.
  __device__ static __forceinline__ void weight_parts(const ModelCtx& c, const double* x, const double* y, double& a, double& f) {
    const double r = y[0] - (c.theta[0] + c.theta[1] * x[0]);
    f = rsqrt(x[0]);
    const double rf = r * f;
    a = -0.5 * (rf * rf);
  }
  __device__ static __forceinline__ double w_const() { return -CPIMH_M_LN_SQRT_2PI; }
};

// ------------------------------------------------------------------ get_stochastic_kinetic (d = 4, d_y = 2)
struct SkModel {
  static constexpr int D = 4;
  static constexpr bool USES_SHARED_U = false;
  static constexpr bool MULTI = true;
  static constexpr bool JUMP_LIST = false;
  static constexpr bool HAS_DTRANS = false;
  static constexpr bool SPLIT_W = false;
  __device__ static double step_const(const ModelCtx&, int) { return 0.0; }
  __device__ static void init(const ModelCtx&, int, double* x) { x[0] = 8.0; x[1] = 8.0; x[2] = 8.0; x[3] = 5.0; }
  __device__ static __forceinline__ double hazards(const double* x, double* h) {      // R/model_stochastic_kinetic.R:27-40
    h[0] = 0.1 * x[3] * x[2];
    h[1] = 0.7 * (5.0 - x[3]);
    h[2] = 0.35 * x[3];
    h[3] = 0.2 * x[0];
    h[4] = 0.1 * x[1] * (x[1] - 1.0) / 2.0;
    h[5] = 0.9 * x[2];
    h[6] = 0.3 * x[0];
    h[7] = 0.1 * x[1];
    double h0 = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++) h0 += h[r];
    return h0;
  }
  __device__ static __forceinline__ int fire(double* x, const double* h, double u_rnd) {     // :58-71, stoichiometry :19-23
    double cdf = 0.0;
    int ev = 0, clamped = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) { cdf += h[r]; ev += (cdf < u_rnd) ? 1 : 0; }
    if (ev > 7) { ev = 7; clamped = 256; }
    switch (ev) {
      case 0: x[2] -= 1.0; x[3] -= 1.0; break;
      case 1: x[2] += 1.0; x[3] += 1.0; break;
      case 2: x[0] += 1.0; break;
      case 3: x[1] += 1.0; break;
      case 4: x[1] -= 2.0; x[2] += 1.0; break;
      case 5: x[1] += 2.0; x[2] -= 1.0; break;
      case 6: x[0] -= 1.0; break;
      default: x[1] -= 1.0; break;
    }
    return clamped;
  }
  __device__ static __forceinline__ void draw(const ModelCtx& c, int t, int i, int m, double& E, double& U) {
    if (c.inj_trans) { E = trans_row(c, t, m)[i]; U = trans_row(c, t, c.sk_M + m)[i]; }
    else { double ua, ub; philox_u2(c.stream, i, t, CPIMH_P_TRANS, m, ua, ub); E = -rng_log(ua); U = ub; }
  }
  // exact Gillespie over one inter-observation interval of 0.1 (:96-129).  Returns 0, or CPIMH_ERR_NOISE_OVERFLOW when
  // the injected event budget is exhausted; bit 8 set if an event index was clamped.
  __device__ static int transition(const ModelCtx& c, int t, int i, double* x, double, uint2) {
    double tcur = 0.0;
    int flags = 0;
    for (int m = 0;; m++) {
      if (c.inj_trans && m >= c.sk_M) return CPIMH_ERR_NOISE_OVERFLOW;
      double h[8], E, U;
      const double h0 = hazards(x, h);
      draw(c, t, i, m, E, U);
      tcur = tcur + (1.0 / h0) * E;            // rexp(n, rate) = exp_rand() * (1 / rate)
      if (!(tcur < 0.1)) return flags;
      flags |= fire(x, h, U * h0);
    }
  }
  // All E particles of one thread as one event stream: when a particle's clock passes 0.1 the lane moves straight on to
  // its next particle, so lanes idle only at the end (event counts per particle-step are heavy-tailed, mean ~1.7).
  __device__ static int transition_multi(const ModelCtx& c, int t, int E, int i0, int stride, double (*xs)[4], const double*, double) {
    int flags = 0, e = 0, m = 0;
    double tcur = 0.0;
    double x[4] = {xs[0][0], xs[0][1], xs[0][2], xs[0][3]};
    while (e < E) {
      if (c.inj_trans && m >= c.sk_M) return CPIMH_ERR_NOISE_OVERFLOW;
      double h[8], Ex, U;
      const double h0 = hazards(x, h);
      draw(c, t, i0 + e * stride, m, Ex, U);
      tcur = tcur + (1.0 / h0) * Ex;
      if (!(tcur < 0.1)) {
        xs[e][0] = x[0]; xs[e][1] = x[1]; xs[e][2] = x[2]; xs[e][3] = x[3];
        e++; m = 0; tcur = 0.0;
        if (e < E) { x[0] = xs[e][0]; x[1] = xs[e][1]; x[2] = xs[e][2]; x[3] = xs[e][3]; }
      } else {
        flags |= fire(x, h, U * h0);
        m++;
      }
    }
    return flags;
  }
  __device__ static double logw(const ModelCtx& c, const double* x, const double* y) {
    const double isd = c.pre[0], log_sd = c.pre[1];       // 1 / sqrt(s2_obs) and log(sqrt(s2_obs)): per-filter constants (host precompute)
    const double* A = c.theta + 1;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      double m = 0.0;
#pragma unroll
      for (int l = 0; l < 4; l++) m += A[j + 2 * l] * x[l];
      const double z = (m - y[j]) * isd;                  // dnorm(m, y_j, sd, log = TRUE), R/model_stochastic_kinetic.R:146-153
      s += -(CPIMH_M_LN_SQRT_2PI + 0.5 * z * z + log_sd);
    }
    return s;
  }
};

// ------------------------------------------------------------------ get_lorenz, thread-per-particle (small d)
template <int D>
__device__ __forceinline__ void lorenz_rhs(double F, const double* x, double* dx) {
#pragma unroll
  for (int n = 0; n < D; n++) {
    const int m1 = (n + D - 1) % D, p1 = (n + 1) % D, m2 = (n + D - 2) % D;
    dx[n] = x[m1] * (x[p1] - x[m2]) - x[n] + F;
  }
}
template <int D>
__device__ __forceinline__ void lorenz_rk4(double F, double h, double* x) {
  double k1[D], k2[D], k3[D], k4[D], tmp[D];
  lorenz_rhs<D>(F, x, k1);
#pragma unroll
  for (int i = 0; i < D; i++) tmp[i] = x[i] + (0.5 * h) * k1[i];
  lorenz_rhs<D>(F, tmp, k2);
#pragma unroll
  for (int i = 0; i < D; i++) tmp[i] = x[i] + (0.5 * h) * k2[i];
  lorenz_rhs<D>(F, tmp, k3);
#pragma unroll
  for (int i = 0; i < D; i++) tmp[i] = x[i] + h * k3[i];
  lorenz_rhs<D>(F, tmp, k4);
#pragma unroll
  for (int i = 0; i < D; i++) x[i] = x[i] + (h / 6.0) * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
}
// Boost.odeint integrate(): controlled Dormand-Prince 5(4), eps_abs = eps_rel = 1e-6
template <int D>
__device__ int lorenz_dopri5(double F, double t0, double t1, double dt, double* x) {
  const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9,
               a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729,
               a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
               c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  const double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
               dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  double k1[D], k2[D], k3[D], k4[D], k5[D], k6[D], k7[D], xt[D], xn[D];
  double t = t0;
  int count = 0;
  lorenz_rhs<D>(F, x, k1);
  while (t1 - t > 2.220446049250313e-16) {
    if ((t + dt) - t1 > 2.220446049250313e-16) dt = t1 - t;
    int trials = 0;
    for (;;) {
#pragma unroll
      for (int i = 0; i < D; i++) xt[i] = x[i] + dt * a21 * k1[i];
      lorenz_rhs<D>(F, xt, k2);
#pragma unroll
      for (int i = 0; i < D; i++) xt[i] = x[i] + dt * (a31 * k1[i] + a32 * k2[i]);
      lorenz_rhs<D>(F, xt, k3);
#pragma unroll
      for (int i = 0; i < D; i++) xt[i] = x[i] + dt * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
      lorenz_rhs<D>(F, xt, k4);
#pragma unroll
      for (int i = 0; i < D; i++) xt[i] = x[i] + dt * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
      lorenz_rhs<D>(F, xt, k5);
#pragma unroll
      for (int i = 0; i < D; i++) xt[i] = x[i] + dt * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
      lorenz_rhs<D>(F, xt, k6);
#pragma unroll
      for (int i = 0; i < D; i++) xn[i] = x[i] + dt * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
      lorenz_rhs<D>(F, xn, k7);
      double err = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) {
        double xe = dt * (dc1 * k1[i] + dc3 * k3[i] + dc4 * k4[i] + dc5 * k5[i] + dc6 * k6[i] + dc7 * k7[i]);
        double e = fabs(xe) / (1e-6 + 1e-6 * (fabs(x[i]) + fabs(dt) * fabs(k1[i])));
        err = fmax(err, e);
      }
      if (err > 1.0) {
        double f = 0.9 * pow(err, -1.0 / 3.0);
        dt *= (f > 0.2 ? f : 0.2);
        if (++trials > 500) return -1;
        continue;
      }
      t += dt;
#pragma unroll
      for (int i = 0; i < D; i++) { x[i] = xn[i]; k1[i] = k7[i]; }
      if (err < 0.5) {
        double e = fmax(err, 0.00032);   // 5^-5
        dt *= 0.9 * pow(e, -1.0 / 5.0);
      }
      break;
    }
    count++;
  }
  return count;
}
template <int D_>
struct LorenzModel {
  static constexpr int D = D_;
  static constexpr bool USES_SHARED_U = false;   // pre[0] = cst of the measurement density
  static constexpr bool MULTI = false;
  static constexpr bool JUMP_LIST = false;
  static constexpr bool HAS_DTRANS = false;
  static constexpr bool SPLIT_W = false;
  __device__ static double step_const(const ModelCtx&, int) { return 0.0; }
  __device__ static void normals(const ModelCtx& c, int i, int t, int purpose, double* z) {
#pragma unroll
    for (int j = 0; 2 * j < D; j++) {
      double ua, ub, z0, z1;
      philox_u2(c.stream, i, t, purpose, j, ua, ub);
      normal2(ua, ub, z0, z1);
      z[2 * j] = z0;
      if (2 * j + 1 < D) z[2 * j + 1] = z1;
    }
  }
  __device__ static void init(const ModelCtx& c, int i, double* x) {
    double z[D];
    if (c.inj_init) {
#pragma unroll
      for (int k = 0; k < D; k++) z[k] = c.inj_init[(size_t)k * c.N + i];
    } else normals(c, i, 0, CPIMH_P_INIT, z);
#pragma unroll
    for (int k = 0; k < D; k++) x[k] = -1.0 + 4.0 * (0.5 * erfc(-z[k] / sqrt(2.0)));   // -1 + 4 pnorm(z)
  }
  __device__ static int transition(const ModelCtx& c, int t, int i, double* x, double, uint2 /*zw*/) {
    const double F = c.theta[0], sx = c.theta[1];
    double z[D];
    if (c.inj_trans) {
#pragma unroll
      for (int k = 0; k < D; k++) z[k] = trans_row(c, t, k)[i];
    } else normals(c, i, t, CPIMH_P_TRANS, z);
    const double t0 = 0.05 * (double)(t - 1), t1 = 0.05 * (double)t;
    if (c.integrator == CPIMH_INTEGRATOR_DOPRI5) lorenz_dopri5<D>(F, t0, t1, 0.05, x);
    else {
      const int ns = c.substeps < 1 ? 1 : c.substeps;
      const double h = (t1 - t0) / ns;
      for (int s = 0; s < ns; s++) lorenz_rk4<D>(F, h, x);
    }
#pragma unroll
    for (int k = 0; k < D; k++) x[k] = x[k] + sx * z[k];
    return 0;
  }
  __device__ static double logw(const ModelCtx& c, const double* x, const double* y) {
    if (isnan(y[0])) return 0.0;
    const double isy = c.pre[1];                          // 1 / sigma_y
    double sq = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) { double r = (x[k] - y[k]) * isy; sq += r * r; }
    return -0.5 * sq + c.pre[0];
  }
};

}  // namespace cpimh
