/* oracle/mvnorm.c -- restatement of src/mvnorm.cpp (fast_dmvnorm* / fast_rmvnorm*) and src/create_A_.cpp.
 * TEST INFRASTRUCTURE (see oracle.h).  The reference delegates the factorisation and the triangular
 * solve to Eigen (RcppEigen, version unpinned, not available here); this file restates the published
 * algorithms: unblocked left-looking Cholesky (Eigen::LLT for small sizes) and column-oriented forward
 * substitution (Eigen's triangular solver for a column-major lower factor).  Normal draws are injected. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define HALF_LOG_2PI_TRUNC 0.9189385   /* src/mvnorm.cpp:73,98,121 -- truncated constant, kept on purpose */

/* lower Cholesky factor of a d x d column-major SPD matrix; returns 0 on success */
int orc_cholesky_lower(const double* cov, int d, double* L) {
  memset(L, 0, sizeof(double) * (size_t)d * d);
  for (int k = 0; k < d; k++) {
    double s = cov[k + (size_t)k * d];
    for (int j = 0; j < k; j++) s -= L[k + (size_t)j * d] * L[k + (size_t)j * d];
    if (!(s > 0)) return 1;
    double lkk = sqrt(s);
    L[k + (size_t)k * d] = lkk;
    for (int i = k + 1; i < d; i++) {
      double t = cov[i + (size_t)k * d];
      for (int j = 0; j < k; j++) t -= L[i + (size_t)j * d] * L[k + (size_t)j * d];
      L[i + (size_t)k * d] = t / lkk;
    }
  }
  return 0;
}

/* src/mvnorm.cpp:113-132 */
void orc_dmvnorm_transpose_cholesky(const double* x, int d, int n, const double* mean, const double* L, double* out) {
  double halflogdet = 0;
  for (int i = 0; i < d; i++) halflogdet += log(L[i + (size_t)i * d]);           /* :120 */
  double cst = -(halflogdet) - (d * HALF_LOG_2PI_TRUNC);                        /* :121 */
  double* r = (double*)malloc(sizeof(double) * (size_t)d);
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < d; i++) r[i] = x[i + (size_t)c * d] - mean[i];          /* :122-126 */
    for (int i = 0; i < d; i++) {                                                /* :127 L^{-1} r, column-oriented */
      r[i] = r[i] / L[i + (size_t)i * d];
      for (int k = i + 1; k < d; k++) r[k] -= r[i] * L[k + (size_t)i * d];
    }
    double sq = 0;
    for (int i = 0; i < d; i++) sq += r[i] * r[i];
    out[c] = -0.5 * sq + cst;                                                    /* :127-130 */
  }
  free(r);
}

/* src/mvnorm.cpp:90-109 */
int orc_dmvnorm_transpose(const double* x, int d, int n, const double* mean, const double* cov, double* out) {
  double* L = (double*)malloc(sizeof(double) * (size_t)d * d);
  if (orc_cholesky_lower(cov, d, L)) { free(L); return 1; }
  orc_dmvnorm_transpose_cholesky(x, d, n, mean, L, out);
  free(L);
  return 0;
}

/* src/mvnorm.cpp:66-84: x is n x d (column-major), mean indexed by column */
int orc_dmvnorm(const double* x, int n, int d, const double* mean, const double* cov, double* out) {
  double* xt = (double*)malloc(sizeof(double) * (size_t)d * n);
  for (int i = 0; i < n; i++) for (int j = 0; j < d; j++) xt[j + (size_t)i * d] = x[i + (size_t)j * n];
  int rc = orc_dmvnorm_transpose(xt, d, n, mean, cov, out);
  free(xt);
  return rc;
}

/* src/mvnorm.cpp:46-62: Y[,i] = rnorm(d) (:52-54, injected as z d x n); Y = L Y (:55); + mean (:56-60) */
void orc_rmvnorm_transpose_cholesky(int n, int d, const double* mean, const double* L, const double* z, double* out) {
  for (int c = 0; c < n; c++)
    for (int i = 0; i < d; i++) {
      double s = 0;
      for (int k = 0; k <= i; k++) s += L[i + (size_t)k * d] * z[k + (size_t)c * d];
      out[i + (size_t)c * d] = s + mean[i];
    }
}

/* src/mvnorm.cpp:26-42 */
int orc_rmvnorm_transpose(int n, int d, const double* mean, const double* cov, const double* z, double* out) {
  double* L = (double*)malloc(sizeof(double) * (size_t)d * d);
  if (orc_cholesky_lower(cov, d, L)) { free(L); return 1; }
  orc_rmvnorm_transpose_cholesky(n, d, mean, L, z, out);
  free(L);
  return 0;
}

/* src/mvnorm.cpp:6-22: Y is n x d filled by column (:12-14), Y = Y U with U = L^T (:10,15), + mean(j) */
int orc_rmvnorm(int n, int d, const double* mean, const double* cov, const double* z, double* out) {
  double* L = (double*)malloc(sizeof(double) * (size_t)d * d);
  if (orc_cholesky_lower(cov, d, L)) { free(L); return 1; }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < d; j++) {
      double s = 0;
      for (int k = 0; k <= j; k++) s += z[i + (size_t)k * n] * L[j + (size_t)k * d];   /* U(k,j) = L(j,k) */
      out[i + (size_t)j * n] = s + mean[j];
    }
  free(L);
  return 0;
}

/* src/create_A_.cpp:6-13 */
void orc_create_A(double alpha, int d, double* A) {
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) A[i + (size_t)j * d] = pow(alpha, 1 + abs(i - j));
}
