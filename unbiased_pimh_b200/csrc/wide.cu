// wide.cu -- the bootstrap filter for WIDE states (Lorenz 96, d = 32 or 64; BASELINE config 5: d = 64, N = 65536,
// T = 500, path storage on): one filter spans the whole GPU.  Same step as pf_fused.cuh (R/CPF.R:35-67) but split at
// the two grid-wide dependencies:
//   wide_propagate      warp per GROUP of G <= 32 consecutive particles: ancestor search, 512-byte coalesced gather of the
//                       ancestor's state, RK4 on shuffles (lorenz_wide.cuh), process noise, measurement density; then
//                       group-local softmax numerators exp(lw - m_group), group sums, the next step's spacings -log(u).
//                       Philox path (two-level CDF): 1024-thread CTAs, one per SM; each CTA's prologue rebuilds the step's
//                       group table (factors exp(m_group - m), CDF over groups, ladder over groups, log Z of the previous
//                       step) in its shared memory from the previous launch's group records -- one launch per time step.
//   wide_weights_scan   one 8-CTA thread-block cluster per filter: the per-PARTICLE normalised CDF (carries exchanged
//                       through distributed shared memory); once after step T for the final path pick, and between steps
//                       only on the injected-uniform parity path.
// State and path storage are the same dense history hist_x[t][i][0..d) (array of rows: a particle's d doubles are
// contiguous, which is also R's d x N column-major layout) + hist_a[t][i].
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "host_common.h"
#include "lorenz_wide.cuh"
#include "wide.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace cpimh {


struct WideParams {
  int N, T, D, B, GX, G, pgrid, substeps;   // GX = number of warp groups (softmax blocks), G <= 32 particles each (any G, not only powers of two)
  unsigned long long seed, first_filter;
  const double* obs;          // [T][D]
  double F, sx, sy, isy, cst;     // isy = 1 / sy
  double* hist_x; int* hist_a;
  double* lw; double* C; double* S;
  // two-level CDF (round 2): per-group inclusive sums instead of per-particle scans; lw is double-buffered by step parity
  double* lw2; int two_level;
  double* pmax; double* psum; double* pv; double* mstep; double* ptot;   // pmax / psum / pv: [2][B][GX], indexed by the parity of the step that wrote them
  size_t maxB;                // stride of the parity dimension (filters allocated)
  double* logz_t; double* logz; double* traj; int* path_index;
  const double* inj_init; const double* inj_trans; const double* inj_resample; const double* inj_path;
  int* anc_rec;
  // exact ancestors (DESIGN.md "Exact ancestors"): the fast pass flags a filter when a uniform lies within rounding distance
  // (margin_rel * sum of the weights) of a CDF knot; the host then reruns the batch with exact = 1 (reference-order CDF)
  int* near; double margin_rel; int exact;
  // all-particle average (R/CPF.R:73-76): normalised weights of step T, walkers, per-CTA partial sums, output [B][T+1][D]
  double* normw; int* walk; double* partial; double* exp_path;
};


// group records (max, sum of numerators, sum of spacings) of step t live in half (t & 1): a fused launch reads step t-1's half
// in its prologue while its faster CTAs already write step t's
__device__ __forceinline__ size_t wide_rec(const WideParams& p, int t, int b) { return ((size_t)(t & 1) * p.maxB + (size_t)b) * (size_t)p.GX; }

__device__ __forceinline__ int g_count_le_w(const double* c, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] <= u) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int g_count_lt_w(const double* c, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] < u) lo = mid + 1; else hi = mid; }
  return lo;
}

// ------------------------------------------------------------------ rinit: -1 + 4 pnorm(z)  (R/model_get_lorenz.R:18-20)
template <int C>
__global__ void wide_init_kernel(WideParams p) {
  constexpr int D = 32 * C;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= p.N) return;
  const Stream s = make_stream(p.seed, p.first_filter + (unsigned long long)b);
  double z[C];
  if (p.inj_init) {
#pragma unroll
    for (int c = 0; c < C; c++) z[c] = p.inj_init[((size_t)b * D + lane * C + c) * p.N + i];
  } else lorenz_wide_normals<C>(s, (unsigned)i, 0, CPIMH_P_INIT, lane, z);
  double* x0 = p.hist_x + ((size_t)b * (p.T + 1)) * p.N * D + (size_t)i * D;
#pragma unroll
  for (int c = 0; c < C; c++) x0[lane * C + c] = -1.0 + 4.0 * (0.5 * erfc(-z[c] / sqrt(2.0)));
}

// ------------------------------------------------------------------ one time step for all particles of all filters
#ifndef WIDE_MINB
#define WIDE_MINB 8
#endif
#ifndef WIDE_UNROLL
#define WIDE_UNROLL 1
#endif
#define WIDE_PRAGMA_(x) _Pragma(#x)
#define WIDE_PRAGMA_UNROLL(n) WIDE_PRAGMA_(unroll n)
// TWO (the two-level CDF: the Philox path when the group table fits shared memory): 1024-thread CTAs, one per SM; every CTA first rebuilds the
// step's group table (factors, inclusive CDF over groups, inclusive ladder over groups) from the previous launch's group
// records in ITS OWN shared memory -- redundant across CTAs (3 * GX doubles read from L2, two scans: ~3 us) but it removes
// the separate one-CTA scan launch (~10 us with its gap while 147 SMs idle) and turns the per-group binary search from 13
// dependent L2 loads into shared-memory loads.
// group table in shared memory: 3 * GX doubles must fit beside the static arrays (227 KB per CTA)
#define WIDE_MAX_GROUPS 8192
template <int C, bool TWO>
__global__ void __launch_bounds__(TWO ? 1024 : 128, TWO ? 1 : WIDE_MINB) wide_propagate_kernel(WideParams p, int t, int ident, int prep_next) {
  constexpr int D = 32 * C;
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int N = p.N, G = p.G;
  const Stream s = make_stream(p.seed, p.first_filter + (unsigned long long)b);
  extern __shared__ __align__(16) unsigned char wide_raw[];
  double* const tab = reinterpret_cast<double*>(wide_raw);      // TWO: [GX] CDF over groups | [GX] ladder over groups | [GX] factors
  __shared__ double scr[34];
  if (TWO && t >= 2) {
    const int GX = p.GX;
    double* gc = tab; double* gp = tab + GX; double* gf = tab + 2 * (size_t)GX;
    const size_t ro = wide_rec(p, t - 1, b);
    // every thread owns a contiguous chunk of <= 8 groups: running sums inside the chunk, one warp scan of the chunk totals,
    // one scan of the 32 warp totals, offsets added in place (both tables in the same passes)
    const int cpt = (GX + (int)blockDim.x - 1) / (int)blockDim.x;
    const int g0 = min(GX, (int)threadIdx.x * cpt), g1 = min(GX, g0 + cpt);
    double m = -INFINITY;
    for (int g = g0; g < g1; g++) m = fmax(m, p.pmax[ro + g]);
    m = block_max(m, scr);
    double aw = 0.0, av = 0.0;
    for (int g = g0; g < g1; g++) {
      const double mc = p.pmax[ro + g];
      const double f = (mc == -INFINITY) ? 0.0 : exp(mc - m);
      gf[g] = f;
      aw += p.psum[ro + g] * f;
      av += p.pv[ro + g];
      gc[g] = aw;
      gp[g] = av;
    }
    const double iw = warp_incl_scan(aw, lane), iv = warp_incl_scan(av, lane);
    double ew = __shfl_up_sync(0xffffffffu, iw, 1), ev = __shfl_up_sync(0xffffffffu, iv, 1);     // exclusive over the warp's threads
    if (lane == 0) { ew = 0.0; ev = 0.0; }
    __shared__ double wtot[2][32];
    if (lane == 31) { wtot[0][warp] = iw; wtot[1][warp] = iv; }
    __syncthreads();
    {
      const double ww = lane < nwarp ? wtot[0][lane] : 0.0, wv = lane < nwarp ? wtot[1][lane] : 0.0;
      const double cw = warp_incl_scan(ww, lane), cv = warp_incl_scan(wv, lane);
      const double ow = (warp > 0 ? __shfl_sync(0xffffffffu, cw, (warp + 31) & 31) : 0.0) + ew;
      const double ov = (warp > 0 ? __shfl_sync(0xffffffffu, cv, (warp + 31) & 31) : 0.0) + ev;
      for (int g = g0; g < g1; g++) { gc[g] += ow; gp[g] += ov; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (blockIdx.x == 0) p.logz_t[(size_t)b * p.T + (t - 2)] = log(gc[GX - 1] / (double)N) + m;   // log Z of step t-1 (R/CPF.R:66)
      if (!ident) scr[32] = gp[GX - 1] - rng_log(philox_u1(s, (unsigned)N, t, CPIMH_P_RESAMPLE, 0));   // the ladder's last spacing
    }
    __syncthreads();
  }
  const double* Cb = p.C + (size_t)b * N;
  const double* Sb = p.S + (size_t)b * N;
  const double* xold = p.hist_x + ((size_t)b * (p.T + 1) + (t - 1)) * N * D;
  double* xnew = p.hist_x + ((size_t)b * (p.T + 1) + t) * N * D;
  int* ha = p.hist_a + ((size_t)b * p.T + (t - 1)) * N;
  const double* y = p.obs + (size_t)(t - 1) * D;
  const bool na = isnan(y[0]);                                             // R/model_get_lorenz.R:39-40
  double yv[C];
#pragma unroll
  for (int c = 0; c < C; c++) yv[c] = y[lane * C + c];
  const double h = 0.05 / (double)p.substeps;
  // TWO: the CDF is kept as inclusive sums GC of the G-particle groups (the prologue above) + the group-scaled numerators
  // lw of the previous step (factor fac per group); likewise the sorted-uniform ladder as group sums GP + the spacings S
  const double* GCb = tab;
  const double* GPb = tab + p.GX;
  const double* facb = tab + 2 * (size_t)p.GX;
  const double* lwold = (TWO ? ((t & 1) ? p.lw2 : p.lw) : p.lw) + (size_t)b * N;           // written by step t-1
  double* lwnew = (TWO ? ((t & 1) ? p.lw : p.lw2) : p.lw) + (size_t)b * N;
  const double scale = ident ? 0.0 : (TWO ? GCb[p.GX - 1] / scr[32] : (p.inj_resample ? Cb[N - 1] : Cb[N - 1] / p.ptot[b]));   // u = sumw e_k  /  (sumw / total) P_k
  // persistent warps: group g = G consecutive particles; the grid covers the machine once and strides over the groups
  for (int grp = blockIdx.x * nwarp + warp; grp < p.GX; grp += gridDim.x * nwarp) {
    const long long base = (long long)grp * G;
    // ancestors of this group's particles: lanes search in parallel (src/multinomial.cpp:25-28 rule)
    const long long k = base + lane;
    const bool mine = lane < G && k < N;
    int a = 0;
    if (mine) {
      if (ident) a = (int)k;
      else if (!TWO) {
        const double u = scale * Sb[k];
        a = g_count_le_w(Cb, N, u);
        if (!p.exact) {                                            // exact = 1: Cb IS the reference-order CDF, nothing to check
          const double mar = p.margin_rel * Cb[N - 1];
          const double cur = a < N ? Cb[a] : INFINITY, below = a > 0 ? Cb[a - 1] : -INFINITY;
          if (cur - u <= mar || u - below < mar) p.near[b] = 1;
        }
        if (a >= N) a = N - 1;
      }
    }
    if (TWO && !ident) {
      // sorted uniform of particle k: (sum of the earlier groups' spacings + prefix inside the group) * scale
      double e = mine ? Sb[k] : 0.0;
      e = warp_incl_scan(e, lane);
      if (mine) {
        const double u = scale * ((grp > 0 ? GPb[grp - 1] : 0.0) + e);
        int gs = g_count_le_w(GCb, p.GX, u);                      // groups whose last knot is <= u
        if (gs >= p.GX) gs = p.GX - 1;
        double acc = gs > 0 ? GCb[gs - 1] : 0.0;
        const double f = facb[gs];
        const long long j0 = (long long)gs * G;
        const int cnt = (int)min((long long)G, (long long)N - j0);
        a = (int)j0 + cnt - 1;
        double below = acc;
        bool hit = false;
        for (int j = 0; j < cnt; j++) {                            // knots inside the group, left to right (src/multinomial.cpp:25-28 rule)
          const double nxt = acc + lwold[j0 + j] * f;
          if (nxt > u) { a = (int)j0 + j; below = acc; acc = nxt; hit = true; break; }
          acc = nxt;
        }
        // the two knots that decided a, against u +- the rounding margin of a CDF summed in another order than the reference's
        const double mar = p.margin_rel * GCb[p.GX - 1];
        if (!hit || acc - u <= mar || (a > 0 && u - below < mar)) p.near[b] = 1;
      }
    }
    if (mine) {
      ha[k] = a;
      if (p.anc_rec) p.anc_rec[((size_t)b * p.T + (t - 1)) * N + k] = a;
    }
    double mylw = -INFINITY;
WIDE_PRAGMA_UNROLL(WIDE_UNROLL)
    for (int j = 0; j < G && base + j < N; j++) {
      const int aj = __shfl_sync(0xffffffffu, a, j);
      const long long kk = base + j;
      double v[C], z[C];
#pragma unroll
      for (int c = 0; c < C; c++) v[c] = xold[(size_t)aj * D + lane * C + c];
      for (int r = 0; r < p.substeps; r++) lorenz_wide_rk4<C>(p.F, h, v, lane);      // R/model_get_lorenz.R:28
      if (p.inj_trans) {
#pragma unroll
        for (int c = 0; c < C; c++) z[c] = p.inj_trans[(((size_t)b * p.T + (t - 1)) * D + lane * C + c) * N + kk];
      } else lorenz_wide_normals<C>(s, (unsigned)kk, t, CPIMH_P_TRANS, lane, z);
      double sq = 0.0;
#pragma unroll
      for (int c = 0; c < C; c++) {
        v[c] = v[c] + p.sx * z[c];                                                      // :29
        xnew[(size_t)kk * D + lane * C + c] = v[c];
        const double r = (v[c] - yv[c]) * p.isy;
        sq += r * r;
      }
      sq = warp_sum(sq);
      const double lw = na ? 0.0 : (-0.5 * sq + p.cst);                                // src/mvnorm.cpp:127-130
      if (lane == j) mylw = lw;
    }
    // Group-local softmax numerators: w = exp(lw - m_g) with m_g the max over THIS group; wide_weights_scan rescales each
    // group by exp(m_g - m) (one exp per group instead of a grid-wide barrier before every exp).  Also the next step's
    // exponential spacings E_k = -log(u_k) (sorted-uniform ladder), so every SM shares the log/exp work.
    const double mg = warp_max(mylw);
    double w = 0.0, v = 0.0;
    if (mine) {
      w = (mg == -INFINITY) ? 0.0 : exp(mylw - mg);                                    // R/CPF.R:62, group-local scaling
      lwnew[k] = w;                       // NOT into C / the buffer being searched: other warps of this launch still read them
      if (prep_next) {
        if (p.inj_resample) v = p.inj_resample[((size_t)b * p.T + t) * N + k];         // row of step t+1
        else v = -rng_log(philox_u1(s, (unsigned)k, t + 1, CPIMH_P_RESAMPLE, 0));
        p.S[(size_t)b * N + k] = v;
        if (p.inj_resample) v = 0.0;
      }
    }
    const double wsum = warp_sum(w), vsum = warp_sum(v);
    if (lane == 0) {
      const size_t wo = wide_rec(p, t, b) + grp;
      p.pmax[wo] = mg;
      p.psum[wo] = wsum;
      p.pv[wo] = vsum;
    }
  }
}

// ------------------------------------------------------------------ normalisation + scans of one step in ONE kernel:
// a thread-block cluster of WIDE_CLUSTER CTAs owns one filter; wide_propagate left block-scaled numerators, per-block
// maxima and sums.  The per-CTA totals are exchanged through distributed shared memory; one cluster barrier replaces a
// kernel boundary.  Per element this kernel only multiplies, divides and scans.
#define WIDE_CLUSTER 8
#define WIDE_TILE 8192       // elements scanned per pass in shared memory (64 KB)
#define WIDE_MAX_BLOCKS_PER_CTA 8192
__global__ void __cluster_dims__(WIDE_CLUSTER, 1, 1) __launch_bounds__(1024, 1)
wide_weights_scan_kernel(WideParams p, int t_done, int prep, int exact) {
  extern __shared__ __align__(16) unsigned char raw[];
  double* buf = reinterpret_cast<double*>(raw);     // [WIDE_TILE]
  double* fac = buf + WIDE_TILE;                    // [blocks of this CTA] exp(m_c - m)
  __shared__ double scr[32];
  __shared__ double tot[2];                         // this CTA's (sum w, sum v), read by the other CTAs of the cluster
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, N = p.N;
  const int G = p.G;                                // softmax block = one warp group of wide_propagate
  const size_t ro = wide_rec(p, t_done, b);
  double m = -INFINITY;
  for (int i = threadIdx.x; i < p.GX; i += blockDim.x) m = fmax(m, p.pmax[ro + i]);
  m = block_max(m, scr);
  // element ranges are aligned to propagate blocks
  const int bper = (p.GX + WIDE_CLUSTER - 1) / WIDE_CLUSTER;
  const int c0 = rank * bper, c1 = min(p.GX, c0 + bper);
  const long long lo = (long long)c0 * G, hi = min((long long)N, (long long)c1 * G);
  double* Cb = p.C + (size_t)b * N;
  double* Sb = p.S + (size_t)b * N;
  const bool ladder = prep && !p.inj_resample;
  double sw = 0.0, sv = 0.0;
  for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    const double mc = p.pmax[ro + c];
    const double f = (mc == -INFINITY) ? 0.0 : exp(mc - m);
    fac[c - c0] = f;
    sw += p.psum[ro + c] * f;
    sv += p.pv[ro + c];
  }
  sw = block_sum(sw, scr);
  sv = block_sum(sv, scr);
  if (threadIdx.x == 0) { tot[0] = sw; tot[1] = sv; }
  cluster.sync();
  double s = 0.0, carry = 0.0, vcarry = 0.0;
  for (int c = 0; c < WIDE_CLUSTER; c++) {
    const double* rt = cluster.map_shared_rank(tot, c);
    const double a = rt[0], v = rt[1];
    if (c < rank) { carry += a; vcarry += v; }
    s += a;
  }
  cluster.sync();                                   // nobody may exit (or overwrite tot) while peers still read it
  if (rank == 0 && threadIdx.x == 0) p.logz_t[(size_t)b * p.T + (t_done - 1)] = log(s / (double)N) + m;   // R/CPF.R:66
  double run = carry / s, vrun = vcarry;
  const double* Wb = p.lw + (size_t)b * N;          // block-scaled numerators written by wide_propagate
  constexpr int EPT = WIDE_TILE / 1024;             // elements per thread and tile: loads are issued together
  for (long long base = lo; base < hi; base += WIDE_TILE) {
    const int n = (int)min((long long)WIDE_TILE, hi - base);
    double r[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; r[e] = q < n ? Wb[base + q] : 0.0; }
#pragma unroll
    for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; if (q < n) buf[q] = (r[e] * fac[(int)(base + q - lo) / G]) / s; }   // normweights (R/CPF.R:63)
    __syncthreads();
    if (p.normw) {
#pragma unroll
      for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; if (q < n) p.normw[(size_t)b * N + base + q] = buf[q]; }
    }
    if (exact) {                                    // the normalised weights themselves: wide_seq_cumsum_kernel sums them left to right
#pragma unroll
      for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; if (q < n) Cb[base + q] = buf[q]; }
    } else {
      block_scan_inplace<false>(buf, n, scr);
#pragma unroll
      for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; if (q < n) Cb[base + q] = buf[q] + run; }
      run += buf[n - 1];
    }
    __syncthreads();
    if (ladder) {
#pragma unroll
      for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; r[e] = q < n ? Sb[base + q] : 0.0; }
#pragma unroll
      for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; if (q < n) buf[q] = r[e]; }
      __syncthreads();
      block_scan_inplace<false>(buf, n, scr);
#pragma unroll
      for (int e = 0; e < EPT; e++) { const int q = threadIdx.x + e * 1024; if (q < n) Sb[base + q] = buf[q] + vrun; }
      vrun += buf[n - 1];
      __syncthreads();
    }
  }
  if (ladder && rank == WIDE_CLUSTER - 1 && threadIdx.x == 0) {
    const Stream st = make_stream(p.seed, p.first_filter + (unsigned long long)b);
    p.ptot[b] = vrun - rng_log(philox_u1(st, (unsigned)N, t_done + 1, CPIMH_P_RESAMPLE, 0));
  }
}

// ------------------------------------------------------------------ exact mode: cumsum(normweights) in the reference's order
// (Rcpp::cumsum, src/multinomial.cpp:17 / the running sum of src/systematic.cpp:13-19): one lane adds left to right, the
// warp stages tiles through shared memory.  ~0.4 ms for N = 65536 -- only filters that flagged a near-knot uniform come here.
__global__ void wide_seq_cumsum_kernel(WideParams p) {
  __shared__ double tile[1024];
  const int b = blockIdx.x, lane = threadIdx.x, N = p.N;
  double* Cb = p.C + (size_t)b * N;
  double acc = 0.0;
  for (int base = 0; base < N; base += 1024) {
    const int n = min(1024, N - base);
    for (int q = lane; q < n; q += 32) tile[q] = Cb[base + q];
    __syncwarp();
    if (lane == 0)
      for (int q = 0; q < n; q++) { acc += tile[q]; tile[q] = acc; }
    __syncwarp();
    for (int q = lane; q < n; q += 32) Cb[base + q] = tile[q];
    __syncwarp();
  }
}

// ------------------------------------------------------------------ final path: systematic(normw, 1, u) + get_path (R/CPF.R:69)
template <int C>
__global__ void wide_path_kernel(WideParams p) {
  constexpr int D = 32 * C;
  const int b = blockIdx.x, lane = threadIdx.x, N = p.N, T = p.T;
  const Stream s = make_stream(p.seed, p.first_filter + (unsigned long long)b);
  const double u = p.inj_path ? p.inj_path[b] : philox_u1(s, 0, T + 1, CPIMH_P_PATH, 0);
  const double* Cb = p.C + (size_t)b * N;
  int k = g_count_lt_w(Cb, N, u);
  if (!p.exact && lane == 0) {                      // `while (csw < u)` (src/systematic.cpp:15): knots k - 1 (< u) and k (>= u)
    const double mar = p.margin_rel * Cb[N - 1];
    const double cur = k < N ? Cb[k] : INFINITY, below = k > 0 ? Cb[k - 1] : -INFINITY;
    if (cur - u < mar || u - below <= mar) p.near[b] = 1;
  }
  if (k >= N) k = N - 1;
  if (lane == 0) {
    p.path_index[b] = k;
    double acc = 0.0;
    for (int t = 0; t < T; t++) acc += p.logz_t[(size_t)b * T + t];
    p.logz[b] = acc;
  }
  int j = k;
  for (int step = T; step >= 0; step--) {
    const double* row = p.hist_x + (((size_t)b * (T + 1) + step) * N + j) * D;
#pragma unroll
    for (int c = 0; c < C; c++) p.traj[((size_t)b * (T + 1) + step) * D + lane * C + c] = row[lane * C + c];
    if (step > 0) j = p.hist_a[((size_t)b * T + (step - 1)) * N + j];
  }
}

// ------------------------------------------------------------------ all_particles = TRUE (R/CPF.R:73-76): the weighted average of
// ALL N paths, exp_path[step] = sum_n normw[n] * x_step[path of particle n].  N walkers start at the leaves and move to their
// ancestors step by step; per step a warp accumulates its particles' 512-byte rows in a fixed order, CTAs reduce through shared
// memory, and a second kernel adds the per-CTA partial sums in CTA order: the result does not depend on scheduling.
#define WIDE_EXP_WARPS 8
template <int C>
__global__ void __launch_bounds__(32 * WIDE_EXP_WARPS) wide_exp_step_kernel(WideParams p, int step) {
  constexpr int D = 32 * C;
  __shared__ double part[WIDE_EXP_WARPS][D];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, N = p.N, T = p.T;
  const double* row0 = p.hist_x + ((size_t)b * (T + 1) + step) * N * D;
  const int* anc = step > 0 ? p.hist_a + ((size_t)b * T + (step - 1)) * N : nullptr;
  const double* w = p.normw + (size_t)b * N;
  int* walk = p.walk + (size_t)b * N;
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; c++) acc[c] = 0.0;
  for (int n = blockIdx.x * WIDE_EXP_WARPS + warp; n < N; n += gridDim.x * WIDE_EXP_WARPS) {
    int j = n;
    if (step < T) { j = lane == 0 ? walk[n] : 0; j = __shfl_sync(0xffffffffu, j, 0); }   // one lane reads: it also overwrites the entry below
    const double wn = w[n];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] += wn * row0[(size_t)j * D + lane * C + c];
    if (anc && lane == 0) walk[n] = anc[j];
  }
#pragma unroll
  for (int c = 0; c < C; c++) part[warp][lane * C + c] = acc[c];
  __syncthreads();
  for (int q = threadIdx.x; q < D; q += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < WIDE_EXP_WARPS; k++) s += part[k][q];
    p.partial[((size_t)b * gridDim.x + blockIdx.x) * D + q] = s;
  }
}
__global__ void wide_exp_reduce_kernel(WideParams p, int step, int nparts) {
  const int b = blockIdx.x, q = threadIdx.x, D = p.D;
  double s = 0.0;
  for (int k = 0; k < nparts; k++) s += p.partial[((size_t)b * nparts + k) * D + q];
  p.exp_path[((size_t)b * (p.T + 1) + step) * D + q] = s;
}

// ================================================================== host side
struct WideFilter {
  cpimh_config cfg;
  int64_t maxB, lastB, launches, bytes;
  int GX, G, pgrid, record_anc;
  // exact ancestors: near-knot flags of the last fast pass, the arguments to rerun it with, and whether that has been settled
  DevBuf near, normw, walk, partial, exp_path;
  int all_particles = 0, exp_parts = 0;
  std::vector<int> near_host;
  uint64_t last_seed = 0, last_first = 0;
  int force_exact = 0, last_exact = 0, resolved = 1;
  std::vector<double> obs_host;
  DevBuf ptot;
  DevBuf obs, hist_x, hist_a, lw, lw2, C, S, pmax, psum, pv, mstep, logz_t, logz, traj, path_index, anc_rec;
  DevBuf inj_init, inj_trans, inj_resample, inj_path;
  bool have_inj[4];
};

static void wide_geometry(const cpimh_config& c, int64_t B, int num_sms, int* G, int* GX) {
  // particles per warp group: the smallest G for which every group of every filter has a resident warp in ONE wave of the
  // fused propagate kernel (32 warps per SM), at most 32 so the per-group ancestor search stays one pass.  Any G in 1..32
  // works (the kernels multiply / divide by G): N = 65536, B = 1 on 148 SMs gives G = 14, 4682 groups on 4736 warps.
  const long long slots = (long long)num_sms * 32;
  long long g = ((long long)c.nparticles * B + slots - 1) / slots;
  if (const char* e = getenv("CPIMH_WIDE_G")) { const int v = atoi(e); if (v >= 1 && v <= 32) g = v; }   // tuning aid only (tools/ab_wide.py)
  const int G_ = (int)(g < 1 ? 1 : (g > 32 ? 32 : g));
  *G = G_;
  *GX = (int)(((long long)c.nparticles + G_ - 1) / G_);
}

int wide_supported(const cpimh_config& c) { return c.model == CPIMH_MODEL_LORENZ && (c.dim == 32 || c.dim == 64); }

int wide_create(const cpimh_config& cfg, int64_t max_filters, WideFilter** out) {
  CPIMH_REQUIRE(cfg.integrator == CPIMH_INTEGRATOR_RK4, "Lorenz 96 with d = %d runs the fixed-step rk4 integrator only (dopri5: d in {4, 8, 16})", cfg.dim);
  CPIMH_REQUIRE(cfg.nparticles >= 1 && cfg.nparticles <= (1 << 20), "wide filter supports 1 <= N <= 2^20 (got %d)", cfg.nparticles);
  CPIMH_REQUIRE(cfg.dim_y == cfg.dim, "dim_y must equal dim for the Lorenz model");
  CPIMH_REQUIRE(!cfg.shuffle, "shuffle = 1 is not available for the wide filter");
  int dev = 0; cudaDeviceProp prop;
  CPIMH_CUDA_CHECK(cudaGetDevice(&dev));
  CPIMH_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  CPIMH_REQUIRE(prop.major >= 10, "libcpimh is built for sm_100a (Blackwell); device %d is sm_%d%d", dev, prop.major, prop.minor);
  WideFilter* h = new WideFilter();
  h->cfg = cfg; h->maxB = max_filters; h->lastB = 0; h->launches = 0; h->bytes = 0; h->record_anc = 0;
  memset(h->have_inj, 0, sizeof(h->have_inj));
  if (const char* e = getenv("CPIMH_WIDE_EXACT")) h->force_exact = atoi(e) ? 1 : 0;      // test aid: every run in the reference-order mode
  wide_geometry(cfg, max_filters, prop.multiProcessorCount, &h->G, &h->GX);
  h->pgrid = prop.multiProcessorCount * WIDE_MINB;          // at least one resident wave of wide_propagate CTAs
  CPIMH_REQUIRE((h->GX + WIDE_CLUSTER - 1) / WIDE_CLUSTER <= WIDE_MAX_BLOCKS_PER_CTA, "N too large for the cluster scan");
  const size_t B = (size_t)max_filters, N = cfg.nparticles, T = cfg.nobs, D = cfg.dim;
  size_t free_b = 0, total_b = 0;
  CPIMH_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
  const double need = (double)B * (T + 1) * N * D * 8.0 + (double)B * T * N * 4.0;
  if (need > 0.9 * (double)free_b) { delete h; CPIMH_REQUIRE(false, "the particle history needs %.1f GB but %.1f GB of HBM is free; lower max_filters", need / 1e9, free_b / 1e9); }
  auto A = [&](DevBuf& b, size_t n) { return b.alloc(n, &h->bytes); };
  int rc = A(h->obs, sizeof(double) * T * D);
  if (!rc) rc = A(h->hist_x, sizeof(double) * B * (T + 1) * N * D);
  if (!rc) rc = A(h->hist_a, sizeof(int) * B * T * N);
  if (!rc) rc = A(h->lw, sizeof(double) * B * N);
  if (!rc) rc = A(h->lw2, sizeof(double) * B * N);
  if (!rc) rc = A(h->C, sizeof(double) * B * N);
  if (!rc) rc = A(h->S, sizeof(double) * B * N);
  if (!rc) rc = A(h->near, sizeof(int) * B);
  if (!rc) rc = A(h->pmax, sizeof(double) * 2 * B * h->GX);
  if (!rc) rc = A(h->psum, sizeof(double) * 2 * B * h->GX);
  if (!rc) rc = A(h->pv, sizeof(double) * 2 * B * h->GX);
  if (!rc) rc = A(h->mstep, sizeof(double) * B);
  if (!rc) rc = A(h->ptot, sizeof(double) * B);
  if (!rc) rc = A(h->logz_t, sizeof(double) * B * T);
  if (!rc) rc = A(h->logz, sizeof(double) * B);
  if (!rc) rc = A(h->traj, sizeof(double) * B * (T + 1) * D);
  if (!rc) rc = A(h->path_index, sizeof(int) * B);
  if (rc) { wide_destroy(h); return rc; }
  *out = h;
  return CPIMH_OK;
}

void wide_destroy(WideFilter* h) {
  if (!h) return;
  DevBuf* all[] = {&h->ptot, &h->obs, &h->hist_x, &h->hist_a, &h->lw, &h->lw2, &h->near, &h->normw, &h->walk, &h->partial, &h->exp_path, &h->C, &h->S, &h->pmax, &h->psum, &h->pv, &h->mstep, &h->logz_t, &h->logz, &h->traj,
                   &h->path_index, &h->anc_rec, &h->inj_init, &h->inj_trans, &h->inj_resample, &h->inj_path};
  for (auto b : all) b->release();
  delete h;
}

int wide_set_observations(WideFilter* h, const double* obs_host) {
  const int T = h->cfg.nobs, D = h->cfg.dim;
  h->obs_host.assign((size_t)T * D, 0.0);
  for (int t = 0; t < T; t++)
    for (int j = 0; j < D; j++) h->obs_host[(size_t)t * D + j] = obs_host[t + (size_t)T * j];
  CPIMH_CUDA_CHECK(cudaMemcpy(h->obs.p, h->obs_host.data(), sizeof(double) * h->obs_host.size(), cudaMemcpyHostToDevice));
  return CPIMH_OK;
}

int wide_set_noise(WideFilter* h, int64_t B, const cpimh_noise* nz) {
  memset(h->have_inj, 0, sizeof(h->have_inj));
  if (!nz) return CPIMH_OK;
  CPIMH_REQUIRE(B >= 1 && B <= h->maxB, "nfilters out of range");
  CPIMH_REQUIRE(!nz->shuffle, "shuffle noise is not available for the wide filter");
  const size_t N = h->cfg.nparticles, T = h->cfg.nobs, D = h->cfg.dim;
  struct Item { const double* src; DevBuf* dst; size_t n; } items[4] = {
      {nz->init, &h->inj_init, (size_t)B * D * N}, {nz->trans, &h->inj_trans, (size_t)B * T * D * N},
      {nz->resample, &h->inj_resample, (size_t)B * T * N}, {nz->path_u, &h->inj_path, (size_t)B}};
  for (int i = 0; i < 4; i++) {
    if (!items[i].src) continue;
    int rc = items[i].dst->alloc(sizeof(double) * items[i].n, &h->bytes);
    if (rc) return rc;
    CPIMH_CUDA_CHECK(cudaMemcpy(items[i].dst->p, items[i].src, sizeof(double) * items[i].n, cudaMemcpyHostToDevice));
    h->have_inj[i] = true;
  }
  return CPIMH_OK;
}

int wide_record_ancestors(WideFilter* h, int on) {
  h->record_anc = on ? 1 : 0;
  if (on) return h->anc_rec.alloc(sizeof(int) * (size_t)h->maxB * h->cfg.nobs * h->cfg.nparticles, &h->bytes);
  return CPIMH_OK;
}

template <int C>
static int wide_run_t(WideFilter* h, const WideParams& p, cudaStream_t s) {
  const int T = p.T, B = p.B;
  const dim3 gi((unsigned)(((long long)p.N * 32 + 255) / 256), (unsigned)B), gp((unsigned)min((p.GX + 3) / 4, max(1, p.pgrid / B)), (unsigned)B), gw(WIDE_CLUSTER, (unsigned)B);
  wide_init_kernel<C><<<gi, 256, 0, s>>>(p);
  h->launches++;
  auto is_ident = [&](int t) { return (t == 1) || isnan(h->obs_host[(size_t)(t - 2) * p.D]); };   // R/CPF.R:38-40
  const size_t wsmem = (WIDE_TILE + WIDE_MAX_BLOCKS_PER_CTA) * sizeof(double);
  CPIMH_CUDA_CHECK(cudaFuncSetAttribute(wide_weights_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
  const size_t fsmem = 3 * (size_t)p.GX * sizeof(double);
  if (p.two_level) CPIMH_CUDA_CHECK(cudaFuncSetAttribute(wide_propagate_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
  const dim3 gf((unsigned)((p.GX + 31) / 32), (unsigned)B);            // two-level: 32 warps per CTA, one group per warp
  for (int t = 1; t <= T; t++) {
    const int ident = is_ident(t);
    if (t >= 2 && !p.two_level) {
      wide_weights_scan_kernel<<<gw, 1024, wsmem, s>>>(p, t - 1, ident ? 0 : 1, p.exact);
      if (p.exact && !ident) { wide_seq_cumsum_kernel<<<B, 32, 0, s>>>(p); h->launches += 1; }
      h->launches += 1;
    }
    const int prep_next = (t < T && !is_ident(t + 1)) ? 1 : 0;
    if (p.two_level) wide_propagate_kernel<C, true><<<gf, 1024, fsmem, s>>>(p, t, ident, prep_next);
    else wide_propagate_kernel<C, false><<<gp, 128, 0, s>>>(p, t, ident, prep_next);
    h->launches++;
  }
  // the last step's weights as a normalised per-particle CDF: the final path pick (R/CPF.R:69) and log Z_T
  WideParams pl = p;
  if (p.two_level && (T & 1) == 0) pl.lw = p.lw2;           // step T wrote the buffer of its parity (its log Z_T comes from this scan)
  pl.normw = h->all_particles ? h->normw.as<double>() : nullptr;
  wide_weights_scan_kernel<<<gw, 1024, wsmem, s>>>(pl, T, 0, p.exact);
  if (p.exact) { wide_seq_cumsum_kernel<<<B, 32, 0, s>>>(p); h->launches += 1; }
  wide_path_kernel<C><<<B, 32, 0, s>>>(p);
  h->launches += 2;
  if (h->all_particles) {
    WideParams pe = p;
    pe.normw = h->normw.as<double>(); pe.walk = h->walk.as<int>(); pe.partial = h->partial.as<double>(); pe.exp_path = h->exp_path.as<double>();
    const dim3 ge((unsigned)h->exp_parts, (unsigned)B);
    for (int step = T; step >= 0; step--) {
      wide_exp_step_kernel<C><<<ge, 32 * WIDE_EXP_WARPS, 0, s>>>(pe, step);
      wide_exp_reduce_kernel<<<B, p.D, 0, s>>>(pe, step, h->exp_parts);
    }
    h->launches += 2 * (T + 1);
  }
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}

static int wide_launch(WideFilter* h, int64_t B, uint64_t seed, uint64_t first_filter, int exact, cudaStream_t s) {
  CPIMH_REQUIRE(B >= 1 && B <= h->maxB, "nfilters %lld outside [1, %lld]", (long long)B, (long long)h->maxB);
  const cpimh_config& c = h->cfg;
  WideParams p;
  memset(&p, 0, sizeof(p));
  p.N = c.nparticles; p.T = c.nobs; p.D = c.dim; p.B = (int)B; p.GX = h->GX; p.G = h->G; p.pgrid = h->pgrid; p.substeps = c.substeps < 1 ? 1 : c.substeps;
  p.seed = seed; p.first_filter = first_filter;
  p.obs = h->obs.as<double>();
  p.F = c.theta[0]; p.sx = c.theta[1]; p.sy = c.theta[2]; p.isy = 1.0 / c.theta[2];
  double hl = 0;
  for (int i = 0; i < c.dim; i++) hl += log(c.theta[2]);                               // src/mvnorm.cpp:120-121
  p.cst = -(hl) - (c.dim * CPIMH_HALF_LOG_2PI_TRUNC);
  p.hist_x = h->hist_x.as<double>(); p.hist_a = h->hist_a.as<int>(); p.lw = h->lw.as<double>(); p.C = h->C.as<double>(); p.S = h->S.as<double>();
  p.lw2 = h->lw2.as<double>();
  // the two-level CDF serves the Philox path; injected sorted uniforms (parity runs) keep the per-particle scans
  p.two_level = (!exact && !h->have_inj[2] && h->GX <= WIDE_MAX_GROUPS) ? 1 : 0;
  p.exact = exact; p.near = h->near.as<int>();
  p.margin_rel = 12.0 * ((double)c.nparticles + 2.0) * 0x1p-53;      // as pf_fused.cuh check_knots: both CDFs are sums of N terms, relative error <= (N - 1) 2^-53 each
  if (!exact) CPIMH_CUDA_CHECK(cudaMemsetAsync(h->near.p, 0, sizeof(int) * (size_t)B, s));
  p.pmax = h->pmax.as<double>(); p.psum = h->psum.as<double>(); p.pv = h->pv.as<double>(); p.mstep = h->mstep.as<double>(); p.ptot = h->ptot.as<double>(); p.maxB = (size_t)h->maxB;
  p.logz_t = h->logz_t.as<double>(); p.logz = h->logz.as<double>(); p.traj = h->traj.as<double>(); p.path_index = h->path_index.as<int>();
  p.inj_init = h->have_inj[0] ? h->inj_init.as<double>() : nullptr;
  p.inj_trans = h->have_inj[1] ? h->inj_trans.as<double>() : nullptr;
  p.inj_resample = h->have_inj[2] ? h->inj_resample.as<double>() : nullptr;
  p.inj_path = h->have_inj[3] ? h->inj_path.as<double>() : nullptr;
  p.anc_rec = h->record_anc ? h->anc_rec.as<int>() : nullptr;
  int rc = c.dim == 64 ? wide_run_t<2>(h, p, s) : wide_run_t<1>(h, p, s);
  if (rc) return rc;
  h->lastB = B; h->last_seed = seed; h->last_first = first_filter; h->last_exact = exact;
  return CPIMH_OK;
}

int wide_run(WideFilter* h, int64_t B, uint64_t seed, uint64_t first_filter, cudaStream_t s) {
  h->resolved = 0;
  h->near_host.assign((size_t)B, 0);
  return wide_launch(h, B, seed, first_filter, h->force_exact, s);
}

// Exact ancestors for one filter across the GPU.  The fast pass records, per filter, whether some uniform came within the
// worst-case rounding distance m = 12 (N + 2) 2^-53 sum(w) of a CDF knot (only then can its ancestor differ from the one the
// reference's left-to-right CDF gives).  With N = 65536 that happens about 2 N^2 m = 0.7 times per STEP -- the bound grows with
// N, and so does the number of knots -- and the only way to settle such a draw is the reference's own serial sum (65536
// dependent additions, 0.3 ms against a 0.05 ms step).  So the reference-order run is a mode of the handle
// (cpimh_pf_exact_mode / CPIMH_WIDE_EXACT=1: every step with cumsum(normweights) by one lane, ~12x slower at N = 65536), and
// the default run REPORTS the flag (cpimh_pf_fetch_exact_repeats) instead of paying for it.
static int wide_resolve(WideFilter* h, cudaStream_t s) {
  if (h->resolved) return CPIMH_OK;
  if (!h->last_exact) {
    CPIMH_CUDA_CHECK(cudaMemcpyAsync(h->near_host.data(), h->near.p, sizeof(int) * (size_t)h->lastB, cudaMemcpyDeviceToHost, s));
    CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  h->resolved = 1;
  return CPIMH_OK;
}
void wide_exact_mode(WideFilter* h, int on) { h->force_exact = on ? 1 : 0; }
int wide_all_particles(WideFilter* h, int on) {
  h->all_particles = on ? 1 : 0;
  if (!on) return CPIMH_OK;
  const size_t B = (size_t)h->maxB, N = h->cfg.nparticles, T = h->cfg.nobs, D = h->cfg.dim;
  int parts = (int)((N + WIDE_EXP_WARPS - 1) / WIDE_EXP_WARPS);
  if (parts > 592) parts = 592;                                    // 4 CTAs of 8 warps per SM
  h->exp_parts = parts;
  int rc = h->normw.alloc(sizeof(double) * B * N, &h->bytes);
  if (!rc) rc = h->walk.alloc(sizeof(int) * B * N, &h->bytes);
  if (!rc) rc = h->partial.alloc(sizeof(double) * B * (size_t)parts * D, &h->bytes);
  if (!rc) rc = h->exp_path.alloc(sizeof(double) * B * (T + 1) * D, &h->bytes);
  return rc;
}
int wide_fetch_exp_paths(WideFilter* h, int64_t B, cudaStream_t s, double* exp_path) {
  CPIMH_REQUIRE(h->all_particles, "call cpimh_pf_all_particles(h, 1) before cpimh_pf_run");
  CPIMH_REQUIRE(B >= 1 && B <= h->lastB, "nfilters %lld exceeds the last run (%lld)", (long long)B, (long long)h->lastB);
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(exp_path, h->exp_path.p, sizeof(double) * (size_t)B * (h->cfg.nobs + 1) * h->cfg.dim, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}
double* wide_device_exp_paths(WideFilter* h) { return h->all_particles ? h->exp_path.as<double>() : nullptr; }
int wide_fetch_exact_repeats(WideFilter* h, int64_t B, cudaStream_t s, int32_t* repeats) {
  CPIMH_REQUIRE(B >= 1 && B <= h->lastB, "nfilters %lld exceeds the last run (%lld)", (long long)B, (long long)h->lastB);
  int rc = wide_resolve(h, s);
  if (rc) return rc;
  for (int64_t b = 0; b < B; b++) repeats[b] = h->near_host[(size_t)b];      // 1: some uniform of this filter came within the rounding margin of a knot (0 in exact mode)
  return CPIMH_OK;
}

int wide_fetch(WideFilter* h, int64_t B, cudaStream_t s, double* logz, double* logz_t, double* trajectory, int32_t* path_index) {
  CPIMH_REQUIRE(B >= 1 && B <= h->lastB, "nfilters %lld exceeds the last run (%lld)", (long long)B, (long long)h->lastB);
  const cpimh_config& c = h->cfg;
  if (logz) CPIMH_CUDA_CHECK(cudaMemcpyAsync(logz, h->logz.p, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, s));
  if (logz_t) CPIMH_CUDA_CHECK(cudaMemcpyAsync(logz_t, h->logz_t.p, sizeof(double) * (size_t)B * c.nobs, cudaMemcpyDeviceToHost, s));
  if (trajectory) CPIMH_CUDA_CHECK(cudaMemcpyAsync(trajectory, h->traj.p, sizeof(double) * (size_t)B * (c.nobs + 1) * c.dim, cudaMemcpyDeviceToHost, s));
  if (path_index) CPIMH_CUDA_CHECK(cudaMemcpyAsync(path_index, h->path_index.p, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}

int wide_fetch_ancestors(WideFilter* h, int64_t b, cudaStream_t s, int32_t* anc) {
  CPIMH_REQUIRE(h->record_anc, "call cpimh_pf_record_ancestors(h, 1) before cpimh_pf_run");
  CPIMH_REQUIRE(b >= 0 && b < h->lastB, "filter index out of range");
  const size_t n = (size_t)h->cfg.nobs * h->cfg.nparticles;
  CPIMH_CUDA_CHECK(cudaMemcpyAsync(anc, h->anc_rec.as<int>() + (size_t)b * n, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  CPIMH_CUDA_CHECK(cudaStreamSynchronize(s));
  return CPIMH_OK;
}

int64_t wide_launch_count(const WideFilter* h) { return h->launches; }
int64_t wide_device_bytes(const WideFilter* h) { return h->bytes; }
void wide_device_outputs(WideFilter* h, double** d_logz, double** d_traj) { if (d_logz) *d_logz = h->logz.as<double>(); if (d_traj) *d_traj = h->traj.as<double>(); }
void wide_effective(const WideFilter* h, cpimh_config* out) { *out = h->cfg; out->tree_capacity = 0; out->threads_per_filter = h->G; out->store_paths = 2; out->reserved = h->GX; }

}  // namespace cpimh
