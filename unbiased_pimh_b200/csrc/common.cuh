// common.cuh -- shared device/host helpers for libcpimh (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/cpimh.h"

namespace cpimh {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
#define CPIMH_CUDA_CHECK(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::cpimh::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                         cudaGetErrorString(_e));                                                \
      return CPIMH_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)
#define CPIMH_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::cpimh::set_error(__VA_ARGS__);           \
      return CPIMH_ERR_INVALID;                  \
    }                                            \
  } while (0)

// ---------------------------------------------------------------- Philox4x32-10 (DESIGN.md "RNG")
// counter = (c0 particle/draw index, c1 time | iteration[0:12) << 20, c2 stream id, c3 purpose | draw << 8 | iteration[12:20) << 24), key = seed
// (time < 2^20, draw < 2^16, iteration < 2^20: the iteration of a chain is the high word of the 64-bit filter id)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// uniforms in the open interval (0,1): ((x >> 12) + 0.5) * 2^-52, exact in fp64 and identical to the oracle's
__device__ __forceinline__ double u64_to_unit(uint32_t lo, uint32_t hi) {
  unsigned long long x = ((unsigned long long)hi << 32) | lo;
  return ((double)(x >> 12) + 0.5) * 0x1p-52;
}
struct Stream {               // identifies one filter's random streams
  uint32_t k0, k1;            // seed
  uint32_t c1hi, c3hi;        // iteration of a chain (high word of the filter id, 20 bits): bits 0..11 << 20 for word 1, bits 12..19 << 24 for word 3
  uint32_t c2;                // stream id
};
__device__ __forceinline__ Stream make_stream(unsigned long long seed, unsigned long long filter_id) {
  Stream s;
  s.k0 = (uint32_t)seed; s.k1 = (uint32_t)(seed >> 32);
  const uint32_t hi = (uint32_t)(filter_id >> 32);
  s.c1hi = hi << 20; s.c3hi = (hi << 12) & 0xff000000u;
  s.c2 = (uint32_t)filter_id;
  return s;
}
__device__ __forceinline__ void philox_u2(const Stream& s, uint32_t i, int t, int purpose, uint32_t draw, double& ua, double& ub) {
  uint4 r = philox4x32_10(make_uint4(i, (uint32_t)t | s.c1hi, s.c2, ((uint32_t)purpose | (draw << 8)) | s.c3hi), make_uint2(s.k0, s.k1));
  ua = u64_to_unit(r.x, r.y);
  ub = u64_to_unit(r.z, r.w);
}
// first uniform (52 bits, words x,y) + the raw second half (words z,w) of the block: models that split the second half into
// two 32-bit uniforms (LevySvModel: Poisson count + first jump) read it in this form
__device__ __forceinline__ void philox_u1raw(const Stream& s, uint32_t i, int t, int purpose, uint32_t draw, double& ua, uint2& zw) {
  uint4 r = philox4x32_10(make_uint4(i, (uint32_t)t | s.c1hi, s.c2, ((uint32_t)purpose | (draw << 8)) | s.c3hi), make_uint2(s.k0, s.k1));
  ua = u64_to_unit(r.x, r.y);
  zw = make_uint2(r.z, r.w);
}
// 32-bit uniform in the open interval (0,1): (x + 0.5) * 2^-32 (exact in fp64; R's own unif_rand() has 32-bit resolution)
__device__ __forceinline__ double u32_to_unit(uint32_t x) { return (__uint2double_rn(x) + 0.5) * 0x1p-32; }
__device__ __forceinline__ double philox_u1(const Stream& s, uint32_t i, int t, int purpose, uint32_t draw) {
  uint4 r = philox4x32_10(make_uint4(i, (uint32_t)t | s.c1hi, s.c2, ((uint32_t)purpose | (draw << 8)) | s.c3hi), make_uint2(s.k0, s.k1));
  return u64_to_unit(r.x, r.y);
}
// ---------------------------------------------------------------- transcendental maps of the RNG specification
// uniforms -> exponential variates (rng_log) and -> Box-Muller angles (rng_sincos2pi).  They belong to the SPECIFICATION of
// the Philox streams (DESIGN.md "RNG"), not to the reference's arithmetic (R's norm_rand / exp_rand cannot be reproduced on
// a GPU): fixed sequences of IEEE-754 double operations, written with explicit round-to-nearest intrinsics so that nvcc
// neither contracts nor reorders them; the CPU oracle runs the same sequence (oracle/rngmath.h), so every uniform maps to the
// same double on both sides.  About half the instructions of log() / sincospi(): no special cases, no table, 1 ulp.
// log(u), u a positive normal double < 2: u = 2^e m, m in (sqrt(1/2), sqrt(2)], s = (m-1)/(m+1), log m = 2 atanh(s).
__device__ __forceinline__ double rng_log(double u) {
  const long long b0 = __double_as_longlong(u);
  int e = (int)(b0 >> 52) - 1023;
  double m = __longlong_as_double((b0 & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll);
  if (m > 0x1.6a09e667f3bcdp+0) { m = __dmul_rn(m, 0.5); e = e + 1; }
  const double f = __dadd_rn(m, -1.0), d = __dadd_rn(m, 1.0);
  const double h = __dmul_rn(d, 0.5);
  const double t = __dadd_rn(1.0, -h);
  double r = __fma_rn(t, __dadd_rn(1.0, t), 1.0);          // 1/h to 1e-2, then three Newton steps
  double eps = __fma_rn(-h, r, 1.0);
  r = __fma_rn(r, eps, r);
  eps = __fma_rn(-h, r, 1.0);
  r = __fma_rn(r, eps, r);
  eps = __fma_rn(-h, r, 1.0);
  r = __fma_rn(r, eps, r);
  const double rd = __dmul_rn(r, 0.5);
  double s = __dmul_rn(f, rd);
  const double rem = __fma_rn(-s, d, f);
  s = __fma_rn(rem, rd, s);                                // s = f/d
  const double z = __dmul_rn(s, s);
  double p = 0x1.8618618618618p-5;                         // 1/21, 1/19, ..., 1/3
  p = __fma_rn(p, z, 0x1.af286bca1af28p-5);
  p = __fma_rn(p, z, 0x1.e1e1e1e1e1e1ep-5);
  p = __fma_rn(p, z, 0x1.1111111111111p-4);
  p = __fma_rn(p, z, 0x1.3b13b13b13b14p-4);
  p = __fma_rn(p, z, 0x1.745d1745d1746p-4);
  p = __fma_rn(p, z, 0x1.c71c71c71c71cp-4);
  p = __fma_rn(p, z, 0x1.2492492492492p-3);
  p = __fma_rn(p, z, 0x1.999999999999ap-3);
  p = __fma_rn(p, z, 0x1.5555555555555p-2);
  const double s2 = __dadd_rn(s, s);
  const double lm = __fma_rn(__dmul_rn(s2, z), p, s2);
  return __fma_rn(__int2double_rn(e), 0x1.62e42fefa39efp-1, lm);
}
// sin(2 pi a), cos(2 pi a), a in [0, 1]: a = q/4 + r, |r| <= 1/8, Taylor series in r^2, quadrant fix-up
template <bool WANT_SIN>
__device__ __forceinline__ void rng_sincos2pi(double a, double& sn, double& cs) {
  const double q = rint(__dmul_rn(4.0, a));
  const double r = __fma_rn(-0.25, q, a);
  const double z = __dmul_rn(r, r);
  const int qi = (int)q & 3;
  double sr = 0.0, cr = 0.0;
  if (WANT_SIN || (qi & 1)) {
    double ps = 0x1.aaec32af93359p-4;
    ps = __fma_rn(ps, z, -0x1.6fadb9f155744p-1);
    ps = __fma_rn(ps, z, 0x1.e8f434d018d63p+1);
    ps = __fma_rn(ps, z, -0x1.e3074fde8871fp+3);
    ps = __fma_rn(ps, z, 0x1.50783487ee782p+5);
    ps = __fma_rn(ps, z, -0x1.32d2cce62bd86p+6);
    ps = __fma_rn(ps, z, 0x1.466bc6775aae2p+6);
    ps = __fma_rn(ps, z, -0x1.4abbce625be53p+5);
    ps = __fma_rn(ps, z, 0x1.921fb54442d18p+2);
    sr = __dmul_rn(r, ps);
  }
  if (WANT_SIN || !(qi & 1)) {
    double pc = -0x1.2a0c591af8314p-5;
    pc = __fma_rn(pc, z, 0x1.20c62c2f2d7f5p-2);
    pc = __fma_rn(pc, z, -0x1.b6e24f44b128fp+0);
    pc = __fma_rn(pc, z, 0x1.f9d38a3763cc3p+2);
    pc = __fma_rn(pc, z, -0x1.a6d1f2a204a8cp+4);
    pc = __fma_rn(pc, z, 0x1.e1f506891babbp+5);
    pc = __fma_rn(pc, z, -0x1.55d3c7e3cbffap+6);
    pc = __fma_rn(pc, z, 0x1.03c1f081b5ac4p+6);
    pc = __fma_rn(pc, z, -0x1.3bd3cc9be45dep+4);
    cr = __fma_rn(z, pc, 1.0);
  }
  sn = (qi & 1) ? cr : sr;
  cs = (qi & 1) ? sr : cr;
  if (qi == 1 || qi == 2) cs = -cs;
  if (qi >= 2) sn = -sn;
}
// Box-Muller pair: r = sqrt(-2 log ua), angle 2 pi ub
__device__ __forceinline__ void normal2(double ua, double ub, double& z0, double& z1) {
  const double r = sqrt(__dmul_rn(-2.0, rng_log(ua)));
  double s, c;
  rng_sincos2pi<true>(ub, s, c);
  z0 = __dmul_rn(r, c);
  z1 = __dmul_rn(r, s);
}
// first normal of the pair only (models that consume one normal per draw): the same value as normal2's z0
__device__ __forceinline__ double normal1(double ua, double ub) {
  const double r = sqrt(__dmul_rn(-2.0, rng_log(ua)));
  double s, c;
  rng_sincos2pi<true>(ub, s, c);
  return __dmul_rn(r, c);
}

// ---------------------------------------------------------------- warp / block primitives (fp64)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
__device__ __forceinline__ int warp_incl_scan_int(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// Block-wide reductions through a 32-entry smem scratch.  All threads of the CTA must call; the
// result is returned to every thread.  `scratch` must not be in use (ends with a barrier).
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : 0.0;
  r = warp_sum(r);
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_max(r);
  __syncthreads();
  return r;
}
// exclusive prefix of one int per thread; returns the exclusive offset, *total = block total
__device__ __forceinline__ int block_excl_scan_int(int v, int* scratch, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = warp_incl_scan_int(v, lane);
  if (lane == 31) scratch[wid] = inc;
  __syncthreads();
  int w = (lane < nw) ? scratch[lane] : 0;
  int winc = warp_incl_scan_int(w, lane);
  int woff = __shfl_sync(0xffffffffu, winc, wid) - __shfl_sync(0xffffffffu, w, wid);
  *total = __shfl_sync(0xffffffffu, winc, nw - 1);
  __syncthreads();
  return woff + inc - v;
}

// In-place inclusive scan of a[0..n) in shared memory by the whole CTA (forward: a[j] = sum_{i<=j};
// REVERSE: a[j] = sum_{i>=j}).  Each warp owns a contiguous range of 32-wide rows, scans them with a
// running carry (conflict-free row accesses + shuffles), warp totals are combined through `scratch`,
// then offsets are added in a second pass.  Ends with a barrier.
template <bool REVERSE>
__device__ __forceinline__ void block_scan_inplace(double* a, int n, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const int rows = (n + 31) >> 5;
  const int rpw = (rows + nw - 1) / nw;
  const int r0 = wid * rpw, r1 = min(rows, r0 + rpw);
  double carry = 0.0;
  for (int r = r0; r < r1; r++) {
    int j = r * 32 + lane;
    int idx = REVERSE ? (n - 1 - j) : j;
    double v = (j < n) ? a[idx] : 0.0;
    v = warp_incl_scan(v, lane) + carry;
    if (j < n) a[idx] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) scratch[wid] = carry;
  __syncthreads();
  double w = (lane < nw) ? scratch[lane] : 0.0;
  double winc = warp_incl_scan(w, lane);
  double off = __shfl_sync(0xffffffffu, winc, wid) - __shfl_sync(0xffffffffu, w, wid);   // exclusive prefix of warp totals
  if (wid > 0) {
    for (int r = r0; r < r1; r++) {
      int j = r * 32 + lane;
      if (j < n) { int idx = REVERSE ? (n - 1 - j) : j; a[idx] += off; }
    }
  }
  __syncthreads();
}

// A shared-memory array addressed as (dynamic window symbol + byte offset) at every use, never as a stored generic pointer:
// the compiler then emits LDS/STS with the offset as an operand, instead of re-deriving a generic address (special-register
// read of the CTA's cluster rank + address arithmetic) each time a spilled pointer is rematerialised.
template <class T>
struct SmemArr {
  int off;          // byte offset in the dynamic shared-memory window
  __device__ __forceinline__ T* ptr() const {
    extern __shared__ __align__(16) unsigned char cpimh_dyn_smem[];
    return reinterpret_cast<T*>(cpimh_dyn_smem + off);
  }
  __device__ __forceinline__ T& operator[](int i) const { return ptr()[i]; }
  __device__ __forceinline__ operator T*() const { return ptr(); }
};

// ---------------------------------------------------------------- padded layout + chunked scans (fused filter)
// The fused filter keeps its two per-particle shared arrays with one pad slot per 16 elements: element i lives at
// pad16(i).  Striped accesses (lane l -> element base + l, base a multiple of 16 per half-warp) stay conflict-free, and so
// do CHUNKED accesses (thread t -> the 8 contiguous elements 8t..8t+7: half-warp start slots 8t + t/2 cover all 16 bank
// pairs), which lets a scan run thread-serially over 8 elements in registers instead of log-stepping through shuffles.
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }
__host__ __device__ inline int padded_len(int n) { return n + (n >> 4) + 1; }

// In-place inclusive scans of a[0..n) (and b[0..n) when TWO) in the padded layout.  Thread t scans chunk t (8 elements)
// serially, chunk totals are scanned across the warp (shuffles) and across warps (scratch), prefixes are added on the way
// out.  More than blockDim chunks: further passes with a running carry.  `scratch` needs 64 doubles.  Ends with a barrier.
template <bool TWO>
__device__ __forceinline__ void block_scan_chunked(double* a, double* b, int n, double* scratch) {
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = (NT + 31) >> 5;
  const int nchunks = (n + 7) >> 3;
  double carry_a = 0.0, carry_b = 0.0;
  for (int c0 = 0; c0 < nchunks; c0 += NT) {
    const int ch = c0 + tid, i0 = ch << 3;
    const int base = pad16(i0);                       // the 8 elements of a chunk never straddle a pad slot
    double va[8], vb[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const bool in = (i0 + j) < n;
      va[j] = in ? a[base + j] : 0.0;
      if (TWO) vb[j] = in ? b[base + j] : 0.0;
    }
#pragma unroll
    for (int j = 1; j < 8; j++) { va[j] += va[j - 1]; if (TWO) vb[j] += vb[j - 1]; }
    const double ta = va[7], ia = warp_incl_scan(ta, lane);
    double tb = 0.0, ib = 0.0;
    if (TWO) { tb = vb[7]; ib = warp_incl_scan(tb, lane); }
    if (lane == 31) { scratch[wid] = ia; if (TWO) scratch[32 + wid] = ib; }
    __syncthreads();
    const double wa = (lane < nw) ? scratch[lane] : 0.0;
    const double wia = warp_incl_scan(wa, lane);
    const double offa = carry_a + (__shfl_sync(0xffffffffu, wia, wid) - __shfl_sync(0xffffffffu, wa, wid)) + (ia - ta);
    carry_a += __shfl_sync(0xffffffffu, wia, nw - 1);
    double offb = 0.0;
    if (TWO) {
      const double wb = (lane < nw) ? scratch[32 + lane] : 0.0;
      const double wib = warp_incl_scan(wb, lane);
      offb = carry_b + (__shfl_sync(0xffffffffu, wib, wid) - __shfl_sync(0xffffffffu, wb, wid)) + (ib - tb);
      carry_b += __shfl_sync(0xffffffffu, wib, nw - 1);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if ((i0 + j) < n) { a[base + j] = va[j] + offa; if (TWO) b[base + j] = vb[j] + offb; }
    }
    __syncthreads();
  }
}
// counts over a padded ascending array: number of entries <= u in [lo, hi) (src/multinomial.cpp:25-27 rule) and number of
// entries < u (src/systematic.cpp:15-18 rule: smallest j with csw_j >= u)
__device__ __forceinline__ int count_le_p(const double* c, int lo, int hi, double u) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c[pad16(mid)] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// two independent bounded searches advanced together: the two chains of dependent shared-memory loads overlap
__device__ __forceinline__ void count_le_p2(const double* c, int lo1, int hi1, double u1, int lo2, int hi2, double u2, int& r1, int& r2) {
  while (lo1 < hi1 || lo2 < hi2) {
    const int m1 = (lo1 + hi1) >> 1, m2 = (lo2 + hi2) >> 1;
    const double v1 = c[pad16(m1)], v2 = c[pad16(m2)];
    if (lo1 < hi1) { if (v1 <= u1) lo1 = m1 + 1; else hi1 = m1; }
    if (lo2 < hi2) { if (v2 <= u2) lo2 = m2 + 1; else hi2 = m2; }
  }
  r1 = lo1; r2 = lo2;
}
__device__ __forceinline__ int count_lt_p(const double* c, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c[pad16(mid)] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}


// ---------------------------------------------------------------- CTA-size independent reductions
// Block sum of f(0) + ... + f(n-1) evaluated as ONE fixed tree whatever the CTA size (a multiple of 32 threads), so that a
// filter returns the same bits from a 32-, 256- or 1024-thread CTA.  The tree: segments of 1024 consecutive terms; inside a
// segment lane l of one warp adds f(1024 g + l), f(1024 g + l + 32), ... in that order, an xor butterfly combines the 32
// lanes; the segment totals are then added left to right.  f(k, kp) gets the term index and its padded slot.  All threads must call; `scratch` needs 32 doubles (n <= 32768)
// and must be free; begins and ends with a barrier.
template <class F>
__device__ __forceinline__ double block_sum_canonical(F f, int n, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nseg = (n + 1023) >> 10;
  __syncthreads();
  for (int g = wid; g < nseg; g += nw) {
    const int k0 = (g << 10) + lane, k1 = min(n, (g + 1) << 10);
    double part = 0.0;
    for (int k = k0, kp = pad16(k0); k < k1; k += 32, kp += 34) part += f(k, kp);     // kp = pad16(k)
    part = warp_sum(part);
    if (lane == 0) scratch[g] = part;
  }
  __syncthreads();
  double r = scratch[0];
  for (int g = 1; g < nseg; g++) r += scratch[g];
  __syncthreads();
  return r;
}
// In-place inclusive prefix sums of a padded array by ONE thread, left to right: r[i] = r[i-1] + x[i], the order of
// Rcpp::cumsum (src/multinomial.cpp:17) and of the running sum in src/systematic.cpp:13-17.
__device__ __forceinline__ void sequential_cumsum_p(double* a, int n) {
  double acc = a[0];
  for (int i = 1; i < n; i++) { acc += a[pad16(i)]; a[pad16(i)] = acc; }
}

#define CPIMH_M_LN_SQRT_2PI 0.918938533204672741780329736406   // R's dnorm(log=TRUE) constant
#define CPIMH_INV_SQRT_10 0.31622776601683794                   // 1 / sqrt(10)
#define CPIMH_LOG_SQRT_10 0x1.26bb1bbb55516p+0                 // log(sqrt(10.0)) as libm rounds it (gss measurement sd)
#define CPIMH_HALF_LOG_2PI_TRUNC 0.9189385                      // src/mvnorm.cpp:73,98,121 (truncated on purpose)

}  // namespace cpimh
