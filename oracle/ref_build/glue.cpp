// glue.cpp -- extern "C" entry points over the reference's own natives, compiled unmodified from /root/reference/src against
// the stand-in Rcpp.h (rcpp_stub/).  TEST INFRASTRUCTURE: tests/test_oracle_vs_reference.py loads oracle/_ref/libref.so to pin
// oracle/*.c; bench.py --impl reference may time it.  Nothing in the product loads it.
// Each function converts plain arrays to the reference's argument types, calls the reference function, and copies the result out.
#include <RcppEigen.h>
#include "systematic.h"       // /root/reference/src/systematic.h
#include "multinomial.h"      // /root/reference/src/multinomial.h
#include "tree.h"             // /root/reference/src/tree.h
using namespace Rcpp;

// declared in the reference's .cpp files only (exported to R through RcppExports.cpp)
IntegerMatrix indexmatching_cpp(NumericVector w1, NumericVector w2, int ndraws);                       // src/coupledresamplingindexmatching.cpp:8
NumericMatrix create_A_(double alpha, int d);                                                          // src/create_A_.cpp:6
NumericMatrix rinitial_SVLevy_cpp(int N, double t, double xi, double w2, double lambda);               // src/SVLevy_functions.cpp:40
NumericMatrix rtransition_SVLevy_cpp(NumericMatrix Xold, double t_1, double t, double xi, double w2, double lambda);   // :58
NumericMatrix dobs_SVLevy_cpp(NumericVector Yt, NumericMatrix Xts, double mu, double beta, bool log);  // :73

#define REF_API extern "C" __attribute__((visibility("default")))

static void push(std::deque<double>& q, const double* p, long n) { for (long i = 0; i < n; i++) q.push_back(p[i]); }
REF_API void ref_clear_queues() { ref_stub::Queues& q = ref_stub::queues(); q.unif.clear(); q.expo.clear(); q.pois.clear(); q.gamma.clear(); q.norm.clear(); q.underflows = 0; ref_stub::oob() = 0; }
REF_API void ref_push_unif(const double* p, long n) { push(ref_stub::queues().unif, p, n); }
REF_API void ref_push_exp(const double* p, long n) { push(ref_stub::queues().expo, p, n); }
REF_API void ref_push_pois(const double* p, long n) { push(ref_stub::queues().pois, p, n); }
REF_API void ref_push_gamma(const double* p, long n) { push(ref_stub::queues().gamma, p, n); }
REF_API long ref_unif_left() { return (long)ref_stub::queues().unif.size(); }
REF_API long ref_oob() { return ref_stub::oob(); }
REF_API long ref_underflows() { return ref_stub::queues().underflows; }

// src/systematic.cpp:7-32
REF_API void ref_systematic_resampling_n(const double* w, int nparticles, int ndraws, double u, int* anc) {
  NumericVector wv(w, (size_t)nparticles);
  IntegerVector a = systematic_resampling_n_(wv, ndraws, u);
  for (int i = 0; i < ndraws; i++) anc[i] = a[i];
}
// src/multinomial.cpp:13-32 (consumes ndraws + ndraws - 1 queued uniforms)
REF_API void ref_multinomial_resampling_n(const double* w, int nparticles, int ndraws, int* anc) {
  NumericVector wv(w, (size_t)nparticles);
  IntegerVector a = multinomial_resampling_n_(wv, ndraws);
  for (int i = 0; i < ndraws; i++) anc[i] = a[i];
}
// src/multinomial.cpp:35-71; anc is ndraws x 2 column-major
REF_API void ref_coupled_multinomial_resampling_n(const double* w1, const double* w2, int nparticles, int ndraws, int* anc) {
  NumericVector a1(w1, (size_t)nparticles), a2(w2, (size_t)nparticles);
  IntegerMatrix a = coupled_multinomial_resampling_n_(a1, a2, ndraws);
  for (int i = 0; i < ndraws; i++) { anc[i] = a(i, 0); anc[ndraws + i] = a(i, 1); }
}
// src/coupledresamplingindexmatching.cpp:8-57
REF_API void ref_indexmatching(const double* w1, const double* w2, int nparticles, int ndraws, int* anc) {
  NumericVector a1(w1, (size_t)nparticles), a2(w2, (size_t)nparticles);
  IntegerMatrix a = indexmatching_cpp(a1, a2, ndraws);
  for (int i = 0; i < ndraws; i++) { anc[i] = a(i, 0); anc[ndraws + i] = a(i, 1); }
}

// src/tree.cpp:6-141 through an opaque handle
REF_API void* ref_tree_new(int N, int M, int dimx) { return new Tree(N, M, dimx); }
REF_API void ref_tree_free(void* t) { delete static_cast<Tree*>(t); }
static NumericMatrix to_matrix(const double* x, int r, int c) { NumericMatrix m(r, c); std::memcpy(m.data(), x, sizeof(double) * (size_t)r * c); return m; }
REF_API void ref_tree_init(void* t, const double* x0) { Tree* T = static_cast<Tree*>(t); T->init(to_matrix(x0, T->dimx, T->N)); }
REF_API void ref_tree_update(void* t, const double* x, const int* a) {
  Tree* T = static_cast<Tree*>(t);
  IntegerVector av(a, (size_t)T->N);
  T->update(to_matrix(x, T->dimx, T->N), av);
}
REF_API int ref_tree_M(void* t) { return static_cast<Tree*>(t)->M; }
REF_API int ref_tree_nsteps(void* t) { return static_cast<Tree*>(t)->nsteps; }
REF_API void ref_tree_arrays(void* t, int* a_star, int* o_star, int* l_star, double* x_star) {
  Tree* T = static_cast<Tree*>(t);
  if (a_star) std::memcpy(a_star, T->a_star.data(), sizeof(int) * (size_t)T->M);
  if (o_star) std::memcpy(o_star, T->o_star.data(), sizeof(int) * (size_t)T->M);
  if (l_star) std::memcpy(l_star, T->l_star.data(), sizeof(int) * (size_t)T->N);
  if (x_star) std::memcpy(x_star, T->x_star.data(), sizeof(double) * (size_t)T->M * T->dimx);
}
REF_API void ref_tree_get_path(void* t, int n, double* path) {
  Tree* T = static_cast<Tree*>(t);
  NumericMatrix p = T->get_path(n);
  std::memcpy(path, p.data(), sizeof(double) * (size_t)T->dimx * (T->nsteps + 1));
}
REF_API void ref_tree_retrieve_xgeneration(void* t, int lag, double* xgen) {
  Tree* T = static_cast<Tree*>(t);
  NumericMatrix g = T->retrieve_xgeneration(lag);
  std::memcpy(xgen, g.data(), sizeof(double) * (size_t)T->dimx * T->N);
}

// src/create_A_.cpp:6-13 (d x d column-major)
REF_API void ref_create_A(double alpha, int d, double* A) { NumericMatrix m = create_A_(alpha, d); std::memcpy(A, m.data(), sizeof(double) * (size_t)d * d); }

// src/SVLevy_functions.cpp:40-81.  X is N x 2 column-major (column 0 = v, column 1 = z), as the reference lays it out.
REF_API void ref_rinitial_SVLevy(int N, double t, double xi, double w2, double lambda, double* X) {
  NumericMatrix m = rinitial_SVLevy_cpp(N, t, xi, w2, lambda);
  std::memcpy(X, m.data(), sizeof(double) * (size_t)N * 2);
}
REF_API void ref_rtransition_SVLevy(const double* Xold, int N, double t_1, double t, double xi, double w2, double lambda, double* Xnew) {
  NumericMatrix m = rtransition_SVLevy_cpp(to_matrix(Xold, N, 2), t_1, t, xi, w2, lambda);
  std::memcpy(Xnew, m.data(), sizeof(double) * (size_t)N * 2);
}
REF_API void ref_dobs_SVLevy(double y, const double* Xts, int N, double mu, double beta, int give_log, double* out) {
  NumericVector Yt(&y, 1);
  NumericMatrix m = dobs_SVLevy_cpp(Yt, to_matrix(Xts, N, 2), mu, beta, give_log != 0);
  std::memcpy(out, m.data(), sizeof(double) * (size_t)N);
}
