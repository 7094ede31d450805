// Rcpp.h -- a stand-in for the small part of the Rcpp API that the reference's hot-path natives use, so that
//   /root/reference/src/{systematic,multinomial,coupledresamplingindexmatching,tree,create_A_,SVLevy_functions}.cpp
// compile UNMODIFIED (byte for byte, from where they lie) into oracle/_ref/libref.so.  TEST INFRASTRUCTURE ONLY: it
// exists to pin the CPU oracle (oracle/*.c) to the reference's own code; nothing in the product links or loads it.
//
// What is reproduced, and how faithfully:
//   * NumericVector / IntegerVector / LogicalVector / NumericMatrix / IntegerMatrix: shared-ownership handles (copying a
//     handle aliases the storage, like an Rcpp vector wrapping one SEXP), zero-initialised, column-major matrices,
//     operator()(i), operator[](i), (i,j), column / row proxies through the placeholder `_`.
//   * sugar: cumsum, sum, exp, vector (+,-,*,/) vector|scalar.  Rcpp evaluates these element by element in plain double
//     arithmetic (Rcpp/sugar/functions/cumsum.h, sum.h: a double accumulator, left to right); so does this stub.
//   * RNG: runif / rexp / rpois / rgamma / rnorm POP values from queues the test fills (ref_stub::push_*): R's Mersenne
//     Twister is the one thing that cannot exist here.  runif(n, a, b) = a + (b - a) u (R nmath/runif.c);
//     rexp(n, rate) = (1 / rate) * e with e a unit exponential (Rcpp/stats/random/rexp.h: scale * exp_rand()).
//   * dnorm(x, mu, sd, log): R nmath/dnorm.c's formula -(M_LN_SQRT_2PI + 0.5 z^2 + log(sd)) -- a restatement (R absent).
//   * RNGScope: empty.  RCPP_MODULE: swallowed (the class body still has to parse, so class_<> is a chainable dummy).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ref_stub {
struct Queues {
  std::deque<double> unif, expo, pois, gamma, norm;
  long underflows = 0;
};
inline Queues& queues() { static Queues q; return q; }
inline double pop(std::deque<double>& q, double fallback) {
  if (q.empty()) { queues().underflows++; return fallback; }
  double v = q.front(); q.pop_front(); return v;
}
inline long& oob() { static long n = 0; return n; }
template <class T> inline T sentinel() { return T(); }
template <> inline double sentinel<double>() { return 1e300; }
}  // namespace ref_stub

namespace Rcpp {

struct Placeholder {};
static const Placeholder _ = Placeholder();

struct RNGScope { RNGScope() {} ~RNGScope() {} };

template <class T>
class Vector {
 public:
  typedef T* iterator;
  typedef const T* const_iterator;
  Vector() : d_(std::make_shared<std::vector<T> >()) {}
  explicit Vector(int n) : d_(std::make_shared<std::vector<T> >((size_t)(n < 0 ? 0 : n), T())) {}
  explicit Vector(long n) : d_(std::make_shared<std::vector<T> >((size_t)(n < 0 ? 0 : n), T())) {}
  explicit Vector(double n) : d_(std::make_shared<std::vector<T> >((size_t)(n < 0 ? 0 : n), T())) {}
  Vector(const T* p, size_t n) : d_(std::make_shared<std::vector<T> >(p, p + n)) {}
  int size() const { return (int)d_->size(); }
  int length() const { return (int)d_->size(); }
  // Rcpp does not bounds-check; the reference reads past the end when sum(w) < u.  The stub makes that
  // case defined: it counts the access (ref_stub::oob()) and hands out a sentinel (+huge) that ends the reference's walks, so
  // a test can tell "the reference was outside its own contract here" from a real answer.
  T& at_(int i) const {
    if (i < 0 || (size_t)i >= d_->size()) { ref_stub::oob()++; static T s; s = ref_stub::sentinel<T>(); return s; }
    return (*d_)[(size_t)i];
  }
  T& operator()(int i) { return at_(i); }
  const T& operator()(int i) const { return at_(i); }
  T& operator[](int i) { return at_(i); }
  const T& operator[](int i) const { return at_(i); }
  iterator begin() { return d_->data(); }
  iterator end() { return d_->data() + d_->size(); }
  const_iterator begin() const { return d_->data(); }
  const_iterator end() const { return d_->data() + d_->size(); }
  T* data() { return d_->data(); }
  const T* data() const { return d_->data(); }
 private:
  std::shared_ptr<std::vector<T> > d_;
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
typedef Vector<int> LogicalVector;     // R logicals are ints

template <class T>
class Matrix;
// a column (or row) of a matrix: assignment copies ELEMENTS (Rcpp::MatrixColumn / MatrixRow semantics)
template <class T>
class MatrixSlice {
 public:
  MatrixSlice(T* base, int n, int stride) : p_(base), n_(n), s_(stride) {}
  MatrixSlice& operator=(const MatrixSlice& o) { for (int i = 0; i < n_; i++) p_[(size_t)i * s_] = o.p_[(size_t)i * o.s_]; return *this; }
  MatrixSlice& operator=(const Vector<T>& v) { for (int i = 0; i < n_; i++) p_[(size_t)i * s_] = v[i]; return *this; }
  int size() const { return n_; }
  T& operator[](int i) { return p_[(size_t)i * s_]; }
  const T& operator[](int i) const { return p_[(size_t)i * s_]; }
 private:
  T* p_; int n_, s_;
};
template <class T>
class Matrix {
 public:
  Matrix() : r_(0), c_(0), v_(0) {}
  Matrix(int r, int c) : r_(r), c_(c), v_(r * c) {}
  int nrow() const { return r_; }
  int ncol() const { return c_; }
  int rows() const { return r_; }
  int cols() const { return c_; }
  T& operator()(int i, int j) { return v_[i + j * r_]; }
  const T& operator()(int i, int j) const { return v_[i + j * r_]; }
  MatrixSlice<T> operator()(Placeholder, int j) { return MatrixSlice<T>(v_.data() + (size_t)j * r_, r_, 1); }
  MatrixSlice<T> operator()(Placeholder, int j) const { return MatrixSlice<T>(const_cast<T*>(v_.data()) + (size_t)j * r_, r_, 1); }
  MatrixSlice<T> operator()(int i, Placeholder) { return MatrixSlice<T>(v_.data() + i, c_, r_); }
  T* data() { return v_.data(); }
  const T* data() const { return v_.data(); }
 private:
  int r_, c_;
  Vector<T> v_;
};
typedef Matrix<double> NumericMatrix;
typedef Matrix<int> IntegerMatrix;

// ---------------------------------------------------------------- sugar (element by element, plain double)
inline NumericVector cumsum(const NumericVector& x) {        // Rcpp/sugar/functions/cumsum.h: result[i] = result[i-1] + x[i]
  NumericVector r(x.size());
  if (x.size() == 0) return r;
  r[0] = x[0];
  for (int i = 1; i < x.size(); i++) r[i] = r[i - 1] + x[i];
  return r;
}
inline double sum(const NumericVector& x) {                  // Rcpp/sugar/functions/sum.h: double accumulator from 0
  double s = 0;
  for (int i = 0; i < x.size(); i++) s += x[i];
  return s;
}
inline NumericVector exp(const NumericVector& x) { NumericVector r(x.size()); for (int i = 0; i < x.size(); i++) r[i] = std::exp(x[i]); return r; }
#define REF_STUB_BINOP(op)                                                                                                             \
  inline NumericVector operator op(const NumericVector& a, const NumericVector& b) { NumericVector r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] op b[i]; return r; } \
  inline NumericVector operator op(const NumericVector& a, double b) { NumericVector r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = a[i] op b; return r; }                  \
  inline NumericVector operator op(double a, const NumericVector& b) { NumericVector r(b.size()); for (int i = 0; i < b.size(); i++) r[i] = a op b[i]; return r; }
REF_STUB_BINOP(+)
REF_STUB_BINOP(-)
REF_STUB_BINOP(*)
REF_STUB_BINOP(/)
#undef REF_STUB_BINOP
inline NumericVector operator-(const NumericVector& a) { NumericVector r(a.size()); for (int i = 0; i < a.size(); i++) r[i] = -a[i]; return r; }
// row proxy = sugar expression (d1logdobs_SVLevy_cpp): first element(s) of the evaluated vector
template <class T>
inline void assign_slice(MatrixSlice<T>& s, const Vector<T>& v) { s = v; }

// ---------------------------------------------------------------- random variates: popped from the injected queues
inline NumericVector runif(int n) { NumericVector r(n); for (int i = 0; i < n; i++) r[i] = ref_stub::pop(ref_stub::queues().unif, 0.5); return r; }
inline NumericVector runif(int n, double a, double b) { NumericVector r(n); for (int i = 0; i < n; i++) r[i] = a + (b - a) * ref_stub::pop(ref_stub::queues().unif, 0.5); return r; }
inline NumericVector rexp(int n, double rate) { const double scale = 1.0 / rate; NumericVector r(n); for (int i = 0; i < n; i++) r[i] = scale * ref_stub::pop(ref_stub::queues().expo, 1.0); return r; }
inline NumericVector rpois(int n, double) { NumericVector r(n); for (int i = 0; i < n; i++) r[i] = ref_stub::pop(ref_stub::queues().pois, 0.0); return r; }
inline NumericVector rgamma(int n, double, double) { NumericVector r(n); for (int i = 0; i < n; i++) r[i] = ref_stub::pop(ref_stub::queues().gamma, 1.0); return r; }
inline NumericVector rnorm(int n) { NumericVector r(n); for (int i = 0; i < n; i++) r[i] = ref_stub::pop(ref_stub::queues().norm, 0.0); return r; }

// ---------------------------------------------------------------- densities
inline NumericVector dnorm(const NumericVector& x, double mu, double sigma, bool give_log) {   // R nmath/dnorm.c
  NumericVector r(x.size());
  for (int i = 0; i < x.size(); i++) {
    const double z = std::fabs((x[i] - mu) / sigma);
    r[i] = give_log ? -(0.918938533204672741780329736406 + 0.5 * z * z + std::log(sigma))
                    : 0.398942280401432677939946059934 * std::exp(-0.5 * z * z) / sigma;
  }
  return r;
}

// ---------------------------------------------------------------- modules: parsed, never run
template <class C>
struct class_ {
  explicit class_(const char*) {}
  template <class A, class B, class D> class_& constructor() { return *this; }
  template <class M> class_& field(const char*, M) { return *this; }
  template <class F> class_& method(const char*, F) { return *this; }
};
#define RCPP_MODULE(name) static void ref_stub_module_##name()

}  // namespace Rcpp

// a matrix row assigned from a sugar expression (SVLevy_functions.cpp d1logdobs): MatrixSlice::operator=(Vector) covers it
