/* oracle/models.c -- restatement of the five shipped models' rinit / rtransition / dmeasurement.
 * TEST INFRASTRUCTURE (see oracle.h).
 *   get_gss                 R/model_get_gss.R:13-46
 *   get_ar(d)               R/model_get_ar.R:13-39, src/create_A_.cpp:6-13, src/mvnorm.cpp:113-132
 *   get_levy_sv             R/model_levy_sv.R:17-66, src/SVLevy_functions.cpp:40-81
 *   get_stochastic_kinetic  R/model_stochastic_kinetic.R:16-157
 *   get_lorenz              R/model_get_lorenz.R:18-50, src/lorenz96.cpp:13-32,65-93 (+ Boost.odeint integrate())
 * State matrices are d x N column-major (R layout).  Noise is component-major [n][N] (see DESIGN.md
 * "noise injection"); when no array is injected it is drawn from the Philox streams of DESIGN.md "RNG",
 * which the CUDA library implements independently to the same specification. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"
#include "rngmath.h"

#define M_LN_SQRT_2PI_FULL 0.918938533204672741780329736406 /* R's dnorm(log=TRUE) constant (nmath/dnorm.c) */
#define TWO_PI 6.283185307179586476925286766559

static inline uint32_t c1_of(uint64_t filter_id, int t) { return (uint32_t)t | ((uint32_t)(filter_id >> 32) << 20); }
/* bits 12..19 of the iteration (high word of the filter id) go to the top byte of counter word 3 (DESIGN.md "RNG") */
static inline uint32_t c3_of(uint64_t filter_id, uint32_t c3) { return c3 | ((((uint32_t)(filter_id >> 32)) << 12) & 0xff000000u); }
static inline void uni2(uint64_t seed, uint64_t filter_id, uint32_t i, int t, int purpose, uint32_t draw, double* ua, double* ub) {
  orc_philox_uniform2(seed, i, c1_of(filter_id, t), (uint32_t)filter_id, c3_of(filter_id, (uint32_t)purpose | (draw << 8)), ua, ub);
}
static inline void normal2(double ua, double ub, double* z0, double* z1) {
  double r = sqrt(-2.0 * orc_rng_log(ua)), s, c;      /* the RNG specification's log / sincos (rngmath.h) */
  orc_rng_sincos2pi(ub, &s, &c);
  *z0 = r * c;
  *z1 = r * s;
}
/* R nmath/dnorm.c, give_log branch */
static inline double dnorm_log(double x, double mu, double sigma) {
  double z = (x - mu) / sigma;
  return -(M_LN_SQRT_2PI_FULL + 0.5 * z * z + log(sigma));
}

int orc_model_n_init(const orc_config* c) {
  switch (c->model) {
    case ORC_MODEL_GSS: return 1;
    case ORC_MODEL_AR: case ORC_MODEL_LORENZ: return c->dim;
    case ORC_MODEL_LEVY_SV: return 3;   /* z0 ~ Gamma, sum_weighted_e, sum_e */
    default: return 0;
  }
}
int orc_model_n_trans(const orc_config* c) {
  switch (c->model) {
    case ORC_MODEL_GSS: return 1;
    case ORC_MODEL_AR: case ORC_MODEL_LORENZ: return c->dim;
    case ORC_MODEL_LEVY_SV: return 2;   /* sum_weighted_e, sum_e (R/model_get_levydriven.R:49-59 form) */
    case ORC_MODEL_SK: return 2 * (c->sk_max_events > 0 ? c->sk_max_events : ORC_SK_MAX_EVENTS_DEFAULT);
    default: return 0;
  }
}

/* ---------------- Philox noise generation (DESIGN.md "RNG"; mirrored in csrc/models.cuh) ---------------- */
/* Poisson count by sequential inversion (recurrence p_k = p_{k-1} rate / k) AND the position of u inside its cell,
 * (u - F_{k-1}) / p_k: uniform on (0,1), independent of k (DESIGN.md "RNG": it is the time of the first jump) */
static int poisson_inv_cell(double rate, double u, double* cell) {
  double p = exp(-rate), F = p, Fprev = 0.0; int k = 0;
  while (u > F && k < 1000) { k++; p = p * rate / k; Fprev = F; F += p; }
  *cell = (u - Fprev) * (1.0 / p);
  return k;
}
/* compound-Poisson sums of the Levy-SV transition for particle i at time t (src/SVLevy_functions.cpp:61-67).  The block whose
 * first half is the resampling draw of slot i carries, in its second half, two 32-bit uniforms: u_z -> jump count and first jump
 * time, u_w -> first jump size; jumps j >= 2 use block (i, t, TRANS, j). */
static void sv_jump_sums(const orc_config* c, uint64_t seed, uint64_t fid, uint32_t i, int t, double* swe, double* se) {
  double xi = c->theta[2], w2 = c->theta[3], lambda = c->theta[4];
  double ua, ub, cell;
  uint32_t z, w;
  orc_philox_uniform1_raw(seed, i, c1_of(fid, t), (uint32_t)fid, c3_of(fid, ORC_P_RESAMPLE), &ua, &z, &w);
  const double uz = ((double)z + 0.5) * 0x1p-32, uw = ((double)w + 0.5) * 0x1p-32;
  int k = poisson_inv_cell(lambda * xi * xi / w2, uz, &cell);
  double a = 0, b = 0, escale = w2 / xi;
  if (k >= 1) {
    double e = -orc_rng_log(uw) * escale;            /* rexp(k, rate=xi/w2) = exp_rand() * (w2/xi) */
    a = exp(-lambda * cell) * e;                     /* exp(-lambda (t - c)), t - c ~ U(0,1) */
    b = e;
  }
  for (int j = 2; j <= k; j++) {
    uni2(seed, fid, i, t, ORC_P_TRANS, (uint32_t)j, &ua, &ub);
    double e = -orc_rng_log(ub) * escale;
    a += exp(-lambda * ua) * e;
    b += e;
  }
  *swe = a; *se = b;
}
static double gamma_mt(uint64_t seed, uint64_t fid, uint32_t i, double shape, double scale) {
  double a = shape < 1.0 ? shape + 1.0 : shape;
  double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
  for (uint32_t m = 0; m < 64; m++) {
    double ua, ub, z, z1, u2, ub2;
    uni2(seed, fid, i, 0, ORC_P_INIT, 2 * m, &ua, &ub);
    normal2(ua, ub, &z, &z1);
    double v = 1.0 + cc * z;
    if (v <= 0) continue;
    v = v * v * v;
    uni2(seed, fid, i, 0, ORC_P_INIT, 2 * m + 1, &u2, &ub2);
    if (log(u2) < 0.5 * z * z + d - d * v + d * log(v)) {
      double g = d * v;
      if (shape < 1.0) g *= pow(ub2, 1.0 / shape);
      return g * scale;
    }
  }
  return shape * scale;
}

void orc_fill_init_noise(const orc_config* c, uint64_t seed, uint64_t fid, double* noise) {
  int N = c->nparticles;
  double ua, ub, z0, z1;
  switch (c->model) {
    case ORC_MODEL_GSS:
      for (int i = 0; i < N; i++) { uni2(seed, fid, i, 0, ORC_P_INIT, 0, &ua, &ub); normal2(ua, ub, &z0, &z1); noise[i] = z0; }
      break;
    case ORC_MODEL_AR: case ORC_MODEL_LORENZ:
      for (int i = 0; i < N; i++)
        for (int j = 0; 2 * j < c->dim; j++) {
          uni2(seed, fid, i, 0, ORC_P_INIT, (uint32_t)j, &ua, &ub); normal2(ua, ub, &z0, &z1);
          noise[(size_t)(2 * j) * N + i] = z0;
          if (2 * j + 1 < c->dim) noise[(size_t)(2 * j + 1) * N + i] = z1;
        }
      break;
    case ORC_MODEL_LEVY_SV: {
      double xi = c->theta[2], w2 = c->theta[3];
      for (int i = 0; i < N; i++) {
        noise[i] = gamma_mt(seed, fid, i, xi * xi / w2, w2 / xi);          /* src/SVLevy_functions.cpp:45 */
        sv_jump_sums(c, seed, fid, i, 0, &noise[(size_t)N + i], &noise[(size_t)2 * N + i]);
      }
      break;
    }
    default: break;
  }
}
void orc_fill_trans_noise(const orc_config* c, uint64_t seed, uint64_t fid, int time, double* noise) {
  int N = c->nparticles;
  double ua, ub, z0, z1;
  switch (c->model) {
    case ORC_MODEL_GSS:
      for (int i = 0; i < N; i++) { uni2(seed, fid, i, time, ORC_P_TRANS, 0, &ua, &ub); normal2(ua, ub, &z0, &z1); noise[i] = z0; }
      break;
    case ORC_MODEL_AR: case ORC_MODEL_LORENZ:
      for (int i = 0; i < N; i++)
        for (int j = 0; 2 * j < c->dim; j++) {
          uni2(seed, fid, i, time, ORC_P_TRANS, (uint32_t)j, &ua, &ub); normal2(ua, ub, &z0, &z1);
          noise[(size_t)(2 * j) * N + i] = z0;
          if (2 * j + 1 < c->dim) noise[(size_t)(2 * j + 1) * N + i] = z1;
        }
      break;
    case ORC_MODEL_LEVY_SV:
      for (int i = 0; i < N; i++) sv_jump_sums(c, seed, fid, i, time, &noise[i], &noise[(size_t)N + i]);
      break;
    case ORC_MODEL_SK: {
      int M = c->sk_max_events > 0 ? c->sk_max_events : ORC_SK_MAX_EVENTS_DEFAULT;
      for (int i = 0; i < N; i++)
        for (int m = 0; m < M; m++) {
          uni2(seed, fid, i, time, ORC_P_TRANS, (uint32_t)m, &ua, &ub);
          noise[(size_t)m * N + i] = -orc_rng_log(ua);
          noise[(size_t)(M + m) * N + i] = ub;
        }
      break;
    }
    default: break;
  }
}

/* ---------------- rinit ---------------- */
void orc_model_rinit(const orc_config* c, const double* noise, double* x) {
  int N = c->nparticles, d = c->dim;
  switch (c->model) {
    case ORC_MODEL_GSS:                                  /* R/model_get_gss.R:14 */
      for (int i = 0; i < N; i++) x[i] = sqrt(2.0) * noise[i];
      break;
    case ORC_MODEL_AR:                                   /* R/model_get_ar.R:14 matrix(rand, nrow=d) */
      for (int i = 0; i < N; i++) for (int k = 0; k < d; k++) x[k + (size_t)i * d] = noise[(size_t)k * N + i];
      break;
    case ORC_MODEL_LORENZ:                               /* R/model_get_lorenz.R:19: -1 + 4 pnorm(rand) */
      for (int i = 0; i < N; i++) for (int k = 0; k < d; k++)
        x[k + (size_t)i * d] = -1.0 + 4.0 * (0.5 * erfc(-noise[(size_t)k * N + i] / sqrt(2.0)));
      break;
    case ORC_MODEL_LEVY_SV: {                            /* src/SVLevy_functions.cpp:40-54 with t = 1 */
      double lambda = c->theta[4];
      int allv = 1, allz = 1;
      for (int i = 0; i < N; i++) {
        double z0 = noise[i], swe = noise[(size_t)N + i], se = noise[(size_t)2 * N + i];
        double z1 = exp(-lambda) * z0 + swe;             /* :51 */
        double v1 = (1 / lambda) * (z0 - z1 + se);       /* :52 */
        x[0 + (size_t)i * 2] = v1; x[1 + (size_t)i * 2] = z1;
        if (!(v1 < 2.220446049250313e-16)) allv = 0;
        if (!(z1 < 2.220446049250313e-16)) allz = 0;
      }
      if (allv) for (int i = 0; i < N; i++) x[(size_t)i * 2] = 2.220446049250313e-16;       /* R/model_levy_sv.R:25 */
      if (allz) for (int i = 0; i < N; i++) x[1 + (size_t)i * 2] = 2.220446049250313e-16;   /* :26 */
      break;
    }
    case ORC_MODEL_SK:                                   /* R/model_stochastic_kinetic.R:43-48: X_0 = (8,8,8,5) */
      for (int i = 0; i < N; i++) { x[(size_t)i * 4] = 8; x[1 + (size_t)i * 4] = 8; x[2 + (size_t)i * 4] = 8; x[3 + (size_t)i * 4] = 5; }
      break;
  }
}

/* ---------------- Lorenz 96 ---------------- */
static void lorenz_rhs(int d, double F, const double* x, double* dx) {    /* src/lorenz96.cpp:21-30, general d */
  for (int n = 0; n < d; n++) {
    int m1 = (n + d - 1) % d, p1 = (n + 1) % d, m2 = (n + d - 2) % d;
    dx[n] = x[m1] * (x[p1] - x[m2]) - x[n] + F;
  }
}
#define LMAX 256
static void rk4_step(int d, double F, double h, double* x) {
  double k1[LMAX], k2[LMAX], k3[LMAX], k4[LMAX], tmp[LMAX];
  lorenz_rhs(d, F, x, k1);
  for (int i = 0; i < d; i++) tmp[i] = x[i] + (0.5 * h) * k1[i];
  lorenz_rhs(d, F, tmp, k2);
  for (int i = 0; i < d; i++) tmp[i] = x[i] + (0.5 * h) * k2[i];
  lorenz_rhs(d, F, tmp, k3);
  for (int i = 0; i < d; i++) tmp[i] = x[i] + h * k3[i];
  lorenz_rhs(d, F, tmp, k4);
  for (int i = 0; i < d; i++) x[i] = x[i] + (h / 6.0) * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
}
/* Boost.odeint integrate() == integrate_adaptive(controlled_runge_kutta<runge_kutta_dopri5>, eps_abs=eps_rel=1e-6)
 * restated from the published algorithm (Boost is not available here => parity unpinned). */
static int dopri5_integrate(int d, double F, double t0, double t1, double dt, double* x) {
  static const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9,
                      a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729,
                      a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
                      c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  const double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640,
               dc5 = c5 - (-92097.0 / 339200), dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  const double eps_abs = 1e-6, eps_rel = 1e-6;
  double k1[LMAX], k2[LMAX], k3[LMAX], k4[LMAX], k5[LMAX], k6[LMAX], k7[LMAX], xt[LMAX], xn[LMAX];
  double t = t0; int count = 0;
  lorenz_rhs(d, F, x, k1);                                  /* fresh stepper: first call evaluates dxdt */
  while (t1 - t > 2.220446049250313e-16) {                  /* less_with_sign(t, t1, dt) */
    if ((t + dt) - t1 > 2.220446049250313e-16) dt = t1 - t;   /* less_with_sign(t1, t + dt, dt) */
    int trials = 0;
    for (;;) {
      for (int i = 0; i < d; i++) xt[i] = x[i] + dt * a21 * k1[i];
      lorenz_rhs(d, F, xt, k2);
      for (int i = 0; i < d; i++) xt[i] = x[i] + dt * (a31 * k1[i] + a32 * k2[i]);
      lorenz_rhs(d, F, xt, k3);
      for (int i = 0; i < d; i++) xt[i] = x[i] + dt * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
      lorenz_rhs(d, F, xt, k4);
      for (int i = 0; i < d; i++) xt[i] = x[i] + dt * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
      lorenz_rhs(d, F, xt, k5);
      for (int i = 0; i < d; i++) xt[i] = x[i] + dt * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
      lorenz_rhs(d, F, xt, k6);
      for (int i = 0; i < d; i++) xn[i] = x[i] + dt * (c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i]);
      lorenz_rhs(d, F, xn, k7);
      double err = 0;
      for (int i = 0; i < d; i++) {
        double xe = dt * (dc1 * k1[i] + dc3 * k3[i] + dc4 * k4[i] + dc5 * k5[i] + dc6 * k6[i] + dc7 * k7[i]);
        double e = fabs(xe) / (eps_abs + eps_rel * (fabs(x[i]) + fabs(dt) * fabs(k1[i])));
        if (e > err) err = e;
      }
      if (err > 1.0) {                                      /* reject: decrease_step, error order 4 */
        double f = 0.9 * pow(err, -1.0 / 3.0);
        dt *= (f > 0.2 ? f : 0.2);
        if (++trials > 500) return -1;
        continue;
      }
      t += dt;                                              /* accept (FSAL: k7 becomes k1) */
      for (int i = 0; i < d; i++) { x[i] = xn[i]; k1[i] = k7[i]; }
      if (err < 0.5) {                                      /* increase_step, stepper order 5, at most x5 */
        double e = err > pow(5.0, -5.0) ? err : pow(5.0, -5.0);
        dt *= 0.9 * pow(e, -1.0 / 5.0);
      }
      break;
    }
    count++;
  }
  return count;
}
void orc_lorenz_step(int d, double F, int integrator, int substeps, double t0, double t1, double dt, double* x) {
  if (integrator == ORC_INTEGRATOR_DOPRI5) { dopri5_integrate(d, F, t0, t1, dt, x); return; }
  if (substeps < 1) substeps = 1;
  double h = (t1 - t0) / substeps;
  for (int s = 0; s < substeps; s++) rk4_step(d, F, h, x);
}

/* ---------------- stochastic kinetic (exact Gillespie) ---------------- */
static const double SK_C[8] = {0.1, 0.7, 0.35, 0.2, 0.1, 0.9, 0.3, 0.1};          /* R/model_stochastic_kinetic.R:24 */
static const int SK_S[4][8] = {{0, 0, 1, 0, 0, 0, -1, 0}, {0, 0, 0, 1, -2, 2, 0, -1},  /* :19-23 rows s1..s4 */
                               {-1, 1, 0, 0, 1, -1, 0, 0}, {-1, 1, 0, 0, 0, 0, 0, 0}};
/* one particle over one inter-observation interval of 0.1 (:96-129); returns #events or -1 on noise overflow.
 * E(m)/U(m): m-th unit exponential / uniform of this particle-step. */
typedef struct { const double* inj; int N, M, i; uint64_t seed, fid; int t; } sk_src;
static inline void sk_draw(const sk_src* s, int m, double* E, double* U) {
  if (s->inj) { *E = s->inj[(size_t)m * s->N + s->i]; *U = s->inj[(size_t)(s->M + m) * s->N + s->i]; }
  else { double ua, ub; uni2(s->seed, s->fid, (uint32_t)s->i, s->t, ORC_P_TRANS, (uint32_t)m, &ua, &ub); *E = -orc_rng_log(ua); *U = ub; }
}
static int sk_particle(double* x, const sk_src* src, int* clamp) {
  const double k = 5;                                                              /* :17 */
  double tcur = 0;
  for (int m = 0;; m++) {
    if (src->inj && m >= src->M) return -1;
    double h[8];
    h[0] = SK_C[0] * x[3] * x[2];                                                  /* :31-38 */
    h[1] = SK_C[1] * (k - x[3]);
    h[2] = SK_C[2] * x[3];
    h[3] = SK_C[3] * x[0];
    h[4] = SK_C[4] * x[1] * (x[1] - 1) / 2;
    h[5] = SK_C[5] * x[2];
    h[6] = SK_C[6] * x[0];
    h[7] = SK_C[7] * x[1];
    long double s = 0; for (int r = 0; r < 8; r++) s += h[r];                      /* colSums: long double */
    double h0 = (double)s, E, U;
    sk_draw(src, m, &E, &U);
    double delta = (1.0 / h0) * E;                                                 /* rexp(n, rate) = exp_rand() * (1/rate), :102 */
    tcur = tcur + delta;                                                           /* :107 */
    if (!(tcur < 0.1)) return m;                                                   /* :108 state frozen */
    double u_rnd = U * h0, cdf = 0; int ev = 0;                                    /* :65-67 */
    for (int r = 0; r < 8; r++) { cdf += h[r]; if (cdf < u_rnd) ev++; }
    if (ev > 7) { ev = 7; if (clamp) (*clamp)++; }                                 /* H8: event index 9 */
    for (int c = 0; c < 4; c++) x[c] += SK_S[c][ev];                               /* :68-69 */
  }
}

/* ---------------- rtransition ---------------- */
/* noise == NULL only for SK in Philox mode (draws on demand from (seed, filter_id)). Returns 0, or <0 on error. */
int orc_model_rtransition_ex(const orc_config* c, int time, const double* noise, uint64_t seed, uint64_t fid, double* x, int* clamp) {
  int N = c->nparticles, d = c->dim;
  switch (c->model) {
    case ORC_MODEL_GSS: {                                 /* R/model_get_gss.R:22-23 */
      double drift = 8 * cos(1.2 * (time - 1));
      for (int i = 0; i < N; i++) {
        double xi = x[i];
        double means = 0.5 * xi + 25 * xi / (1 + xi * xi) + drift;
        x[i] = means + sqrt(1.0) * noise[i];
      }
      return 0;
    }
    case ORC_MODEL_AR: {                                  /* R/model_get_ar.R:21: A %*% x + rand */
      double* A = (double*)malloc(sizeof(double) * (size_t)d * d);
      double* y = (double*)malloc(sizeof(double) * (size_t)d);
      orc_create_A(c->theta[0], d, A);
      for (int i = 0; i < N; i++) {
        for (int r = 0; r < d; r++) { double s = 0; for (int l = 0; l < d; l++) s += A[r + (size_t)l * d] * x[l + (size_t)i * d]; y[r] = s; }
        for (int r = 0; r < d; r++) x[r + (size_t)i * d] = y[r] + noise[(size_t)r * N + i];
      }
      free(A); free(y);
      return 0;
    }
    case ORC_MODEL_LEVY_SV: {                             /* src/SVLevy_functions.cpp:57-70 */
      double lambda = c->theta[4];
      for (int i = 0; i < N; i++) {
        double zold = x[1 + (size_t)i * 2];
        double znew = exp(-lambda) * zold + noise[i];                  /* :66 */
        double vnew = (1 / lambda) * (zold - znew + noise[(size_t)N + i]);  /* :67 */
        x[(size_t)i * 2] = vnew; x[1 + (size_t)i * 2] = znew;
      }
      return 0;
    }
    case ORC_MODEL_SK: {
      int M = c->sk_max_events > 0 ? c->sk_max_events : ORC_SK_MAX_EVENTS_DEFAULT;
      for (int i = 0; i < N; i++) {
        sk_src s = {noise, N, M, i, seed, fid, time};
        if (sk_particle(x + (size_t)i * 4, &s, clamp) < 0) return -2;
      }
      return 0;
    }
    case ORC_MODEL_LORENZ: {                              /* R/model_get_lorenz.R:26-31 */
      for (int i = 0; i < N; i++) {
        orc_lorenz_step(d, c->theta[0], c->integrator, c->substeps, 0.05 * (time - 1), 0.05 * time, 0.05, x + (size_t)i * d);
        for (int r = 0; r < d; r++) x[r + (size_t)i * d] = x[r + (size_t)i * d] + c->theta[1] * noise[(size_t)r * N + i];
      }
      return 0;
    }
  }
  return -1;
}
int orc_model_rtransition(const orc_config* c, int time, const double* noise, double* x) {
  return orc_model_rtransition_ex(c, time, noise, 0, 0, x, 0);
}

/* ---------------- dmeasurement ---------------- */
void orc_model_dmeasurement(const orc_config* c, const double* x, const double* y, double* logw) {
  int N = c->nparticles, d = c->dim;
  switch (c->model) {
    case ORC_MODEL_GSS:                                   /* R/model_get_gss.R:36 */
      for (int i = 0; i < N; i++) logw[i] = dnorm_log(y[0], (x[i] * x[i]) / 20, sqrt(10.0));
      break;
    case ORC_MODEL_AR: case ORC_MODEL_LORENZ: {           /* R/model_get_ar.R:37-38 ; R/model_get_lorenz.R:38-44 */
      double sy = c->model == ORC_MODEL_AR ? c->theta[1] : c->theta[2];
      if (c->model == ORC_MODEL_LORENZ && isnan(y[0])) { for (int i = 0; i < N; i++) logw[i] = 0; break; }
      double* L = (double*)calloc((size_t)d * d, sizeof(double));
      for (int k = 0; k < d; k++) L[k + (size_t)k * d] = sy;          /* theta[2] * diag(d) */
      orc_dmvnorm_transpose_cholesky(x, d, N, y, L, logw);
      free(L);
      break;
    }
    case ORC_MODEL_LEVY_SV: {                             /* src/SVLevy_functions.cpp:73-81 */
      double mu = c->theta[0], beta = c->theta[1];
      for (int i = 0; i < N; i++) { double v = x[(size_t)i * 2]; logw[i] = dnorm_log(y[0], mu + beta * v, sqrt(v)); }
      break;
    }
    case ORC_MODEL_SK: {                                  /* R/model_stochastic_kinetic.R:146-153 */
      double s2 = c->theta[0], sd = sqrt(s2);
      const double* A = c->theta + 1;                     /* 2 x 4 column-major */
      for (int i = 0; i < N; i++) {
        long double s = 0;                                /* colSums: long double */
        for (int j = 0; j < 2; j++) {
          double m = 0;
          for (int l = 0; l < 4; l++) m += A[j + 2 * l] * x[l + (size_t)i * 4];
          s += dnorm_log(m, y[j], sd);
        }
        logw[i] = (double)s;
      }
      break;
    }
  }
}
