// pf_fused.cuh -- the fused bootstrap-particle-filter kernel: one CTA owns one filter for all T steps.
//
// One iteration of the loop below is one iteration of the reference's hot loop R/CPF.R:35-67
// (== R/particlefilter.R:22-38): multinomial resampling (src/multinomial.cpp:13-32) -> gather ->
// rtransition -> dmeasurement -> weight normalisation + log Z increment (R/CPF.R:61-66) ->
// path storage (Tree$update, src/tree.cpp:88-99).  Weights/CDF and sorted uniforms/log-weights stay in shared
// memory for the whole filter.
//
// Path storage has two layouts (PfParams::path_mode):
//   PATH_DENSE  the particle history itself: hist_x[t][q][i] (SoA) and hist_a[t][i].  Step t gathers from
//               hist_x[t-1] and writes hist_x[t], so the state double buffer and the path store are the same
//               bytes; nothing is pruned, get_path() is a walk over hist_a.  This is the B200 layout
//               (180 GB of HBM; (T+1) N (8d+4) bytes per resident CTA).
//   PATH_TREE   the reference's Jacob-Murray-Rubenthaler tree (a_star/o_star/x_star/l_star, prune + insert into
//               the first free slots) with leaf slots, parent slots, offspring counters and the free-slot
//               bitmap in shared memory; O(N log N) nodes, for histories that do not fit in HBM.
#pragma once
#include "common.cuh"
#include "models.cuh"
#ifndef CPIMH_EXP
#define CPIMH_EXP 0      // A/B switches of tools/build_variant.sh (0 in the product)
#endif

namespace cpimh {

enum { PATH_NONE = 0, PATH_DENSE = 1, PATH_TREE = 2 };
#define CPIMH_MAX_E 32        // particles per thread (N <= CPIMH_MAX_E * blockDim.x)

// byte offsets of the shared-memory arrays inside the dynamic window (filled on the host by pf_layout, so that the kernel
// re-derives a pointer with one add from the constant bank instead of recomputing the carve arithmetic)
struct PfLayout { int S, theta, pre, dscr, dscr2, leaf, bpar, off32, bits, coarse, iscr, own; };

struct PfParams {
  PfLayout lay;
  int N, T, dy, Npad;
  int M;                       // tree capacity per filter (multiple of 32)
  long long B;                 // filters in this launch
  long long b0;                // index of the first filter of this launch (outputs and streams are indexed b0 + ...)
  int shuffle, path_mode;
  unsigned long long seed, first_filter;
  const unsigned long long* stream_ids;   // [B] explicit Philox stream id per filter (chain driver), or null: first_filter + b
  const double* obs;           // [T][dy] row-major (device)
  const double* theta;         // [CPIMH_MAX_THETA]
  const double* pre;           // [npre]
  int npre;
  int integrator, substeps;
  // workspaces (indexed by workspace slot = blockIdx.x)
  double* xbuf;                // PATH_NONE/TREE: [ws][2][D][Npad]
  double* hist_x;              // PATH_DENSE:     [ws][T+1][D][Npad]
  int* hist_a;                 // PATH_DENSE:     [ws][T][Npad]
  int* ascr;                   // [ws][Npad] this step's ancestors when the path store is not dense
  double* wscr;                // [ws][Npad] last step's unnormalised weights (exact-order repeat of a resampling step)
  int* exact_repeats;          // [B] or null: number of steps a filter repeated with sequential-order sums
  uint4* jq;                   // [ws][pf_jq_stride(Npad)] jump-list models: per-warp queues of the particles that jump (L2-resident scratch)
  int* a_star;                 // PATH_TREE:      [ws][M]
  int* o_star;                 // PATH_TREE:      [ws][M]
  double* x_star;              // PATH_TREE:      [ws][D][M]
  // outputs (indexed by filter)
  double* logz;                // [B]
  double* logz_t;              // [B][T] or null
  double* traj;                // [B][T+1][D] or null
  int* path_index;             // [B] or null
  int* status;                 // [B]
  double* normw;               // [B][N] or null (final normalised weights)
  int* leaf_out;               // [ws][N] or null (PATH_TREE: final leaf slots)
  int* anc_rec;                // [B][T][N] or null
  double* exp_path;            // [B][T+1][D] or null: weighted average of all N paths (all_particles = TRUE, R/CPF.R:73-76)
  // injected noise (device) or null
  const double* inj_init; const double* inj_trans; const double* inj_resample; const double* inj_shuffle; const double* inj_path;
  int n_init, n_trans, sk_M;
};

struct PfSmem {
  SmemArr<double> C;        // [padded_len(Npad)] weights, then their inclusive CDF          (element i at pad16(i))
  SmemArr<double> S;        // [padded_len(Npad)] exponential spacings -> prefix sums -> log-weights (element i at pad16(i))
  SmemArr<int> leaf;        // [Npad] l_star                                   (PATH_TREE, or shuffle scratch)
  SmemArr<int> bpar;        // [Npad] ancestors, then parent slots b = l_star[a] (PATH_TREE, or shuffle scratch)
  SmemArr<unsigned> off32;  // [Npad/2] 16-bit offspring counters              (PATH_TREE)
  SmemArr<unsigned> bits;   // [M/32] free-slot bitmap (1 = free)              (PATH_TREE)
  SmemArr<double> theta;    // [CPIMH_MAX_THETA]
  SmemArr<double> pre;      // [npre]
  SmemArr<double> dscr;     // [64]
  SmemArr<double> dscr2;    // [4]   E_N of the ladder (written by thread 0 before the scans), sum of the last weights, knot margin, running log Z
  SmemArr<int> coarse;      // [Npad/32 + 1] ancestor of every 32nd sorted uniform (bounds for the per-particle search)
  SmemArr<double> own;      // [16]  the chain kernel's per-replicate scalars (kept out of registers across run_filter)
  SmemArr<int> iscr;        // [36]  (32 scan scratch, [32] status flag, [33] work item, [34] exact-order repeats, [35] near-knot flag)
};
// layout of the dynamic shared-memory window; returns its size in bytes
__host__ __device__ inline size_t pf_layout(int Npad, int M, int npre, int path_mode, int shuffle, PfLayout* L) {
  size_t b = 0;
  PfLayout l;
  b += (size_t)padded_len(Npad) * 8;         // C at offset 0 (padded layout, common.cuh pad16)
  l.S = (int)b; b += (size_t)padded_len(Npad) * 8;
  l.theta = (int)b; b += (size_t)CPIMH_MAX_THETA * 8;
  l.pre = (int)b; b += (size_t)(npre + (npre & 1)) * 8;
  l.dscr = (int)b; b += 64 * 8;
  l.dscr2 = (int)b; b += 4 * 8;
  l.leaf = l.bpar = l.off32 = l.bits = -1;
  if (path_mode == PATH_TREE) {
    l.leaf = (int)b; b += (size_t)Npad * 4;
    l.bpar = (int)b; b += (size_t)Npad * 4;
    l.off32 = (int)b; b += (size_t)(Npad / 2) * 4;
    l.bits = (int)b; b += (size_t)(M / 32) * 4;
  } else if (shuffle) {
    l.bpar = (int)b; b += (size_t)Npad * 4;  // ancestors to permute
  }
  l.coarse = (int)b; b += (size_t)(Npad / 32 + 4) * 4;
  l.iscr = (int)b; b += 36 * 4;
  b = (b + 7) & ~(size_t)7;
  l.own = (int)b; b += 16 * 8;               // chain-owner state (chains.cuh), untouched by run_filter
  if (L) *L = l;
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t pf_smem_bytes(int Npad, int M, int npre, int path_mode, int shuffle) {
  return pf_layout(Npad, M, npre, path_mode, shuffle, nullptr);
}
__device__ __forceinline__ PfSmem pf_carve(const PfLayout& l) {
  PfSmem s;
  s.C.off = 0; s.S.off = l.S; s.theta.off = l.theta; s.pre.off = l.pre; s.dscr.off = l.dscr; s.dscr2.off = l.dscr2;
  s.leaf.off = l.leaf; s.bpar.off = l.bpar; s.off32.off = l.off32; s.bits.off = l.bits;   // -1 when absent (never dereferenced then)
  s.coarse.off = l.coarse; s.iscr.off = l.iscr; s.own.off = l.own;
  return s;
}

// Where one filter's results go.  Only the three outputs that the chain kernel redirects are pointers; everything else is
// addressed at its (rare) use as PfParams array + filter index `fb`, which costs no registers in the hot loop (the arrays are
// kernel parameters in the constant bank; a null array switches the output off; fb < 0 means "no per-filter arrays").
// entries (16 bytes) of jump-queue scratch per workspace: 32 ceil(N / threads) per warp, threads ceil(N / threads) <= N + threads in all
__host__ __device__ inline size_t pf_jq_stride(int Npad) { return (size_t)2 * Npad + 512; }

struct FilterOut {
  double* logz; double* traj; int* status;
};

__device__ __forceinline__ void set_free_bit(unsigned* bits, int slot) { atomicOr(&bits[slot >> 5], 1u << (slot & 31)); }

// Runs one whole filter with the calling CTA.  `ws` selects the HBM workspace, `fb` the filter's row in the per-filter
// output arrays and in the injected noise arrays of PfParams (ignored when they are null).  All threads of the CTA must call.
// LEAN (the production configuration: dense path store, Philox noise, no shuffle, no ancestor recording) compiles the tree, the
// injection pointers and the shuffle out of the kernel: their registers are what the hot loops otherwise spill.
template <class Model, bool LEAN = false>
__device__ void run_filter(const PfParams& p, const PfSmem& sm, long long ws, const Stream& stream, long long fb, const FilterOut& out) {
  constexpr int D = Model::D;
  const int tid = threadIdx.x, NT = blockDim.x;
  // striped particle k = tid + e NT lives at padded slot pad16(k) = kp0 + e NTP (NT is a multiple of 32)
  const int kp0 = pad16(tid), NTP = NT + (NT >> 4);
  const int N = p.N, T = p.T, Npad = p.Npad, M = p.M, W = M >> 5;
  const bool tree = !LEAN && p.path_mode == PATH_TREE, dense = LEAN || p.path_mode == PATH_DENSE;
  const size_t xstride = (size_t)D * Npad;
  double* xold;                // state at time t-1 (SoA [D][Npad])
  double* xnew;
  int* hist_a = nullptr;
  if (dense) {
    xold = p.hist_x + (size_t)ws * (T + 1) * xstride;
    xnew = xold + xstride;
    hist_a = p.hist_a + (size_t)ws * T * Npad;
  } else {
    xold = p.xbuf + (size_t)ws * 2 * xstride;
    xnew = xold + xstride;
  }
  int* a_star = tree ? p.a_star + (size_t)ws * M : nullptr;
  int* o_star = tree ? p.o_star + (size_t)ws * M : nullptr;
  double* x_star = tree ? p.x_star + (size_t)ws * D * M : nullptr;

  ModelCtx c;
  c.theta = sm.theta; c.pre = sm.pre; c.stream = stream; c.N = N; c.T = T;
  c.inj_init = (!LEAN && p.inj_init) ? p.inj_init + (size_t)fb * p.n_init * N : nullptr;
  c.inj_trans = (!LEAN && p.inj_trans) ? p.inj_trans + (size_t)fb * T * p.n_trans * N : nullptr;
  c.n_trans = p.n_trans; c.sk_M = p.sk_M; c.integrator = p.integrator; c.substeps = p.substeps;
  const bool inj_res = !LEAN && p.inj_resample != nullptr;      // injected arrays are addressed at their use (no pointers held across the loop)
  const bool need_shared_u = Model::USES_SHARED_U && !c.inj_trans;
  // status / clamp flags are rare events: they go straight to shared memory (no register held across the filter)
  auto flag_status = [&](int f) { if (f) atomicOr(&sm.iscr[32], f); };
  const bool jump_list = Model::JUMP_LIST && (N + NT - 1) / NT <= 16;   // tokens of <= 16 own particles fit two 64-bit registers

  // ---------------- initialisation: R/CPF.R:17-28, Tree::Tree + Tree::init (src/tree.cpp:6-34)
  if (tid == 0) { sm.iscr[32] = 0; sm.iscr[34] = 0; sm.C[pad16(N)] = INFINITY; }   // status flags, exact-order repeats, CDF sentinel (check_knots)
  if (tree) {
    for (int j = tid; j < M; j += NT) o_star[j] = 0;
    for (int w = tid; w < W; w += NT) {
      int lo = w << 5;
      sm.bits[w] = (lo + 32 <= N) ? 0u : (lo >= N ? 0xffffffffu : (0xffffffffu << (N - lo)));
    }
  }
  for (int i = tid; i < N; i += NT) {
    double x[D];
    Model::init(c, i, x);
#pragma unroll
    for (int k = 0; k < D; k++) xold[(size_t)k * Npad + i] = x[k];
    if (tree) {
#pragma unroll
      for (int k = 0; k < D; k++) x_star[(size_t)k * M + i] = x[k];
      a_star[i] = -2;
      sm.leaf[i] = i;
    }
    sm.C[pad16(i)] = 1.0 / N;
  }
  if (tid == 0) sm.dscr2[3] = 0.0;                       // sum of the log Z increments (thread 0 only)
  __syncthreads();

  // ---------------- hot loop: R/CPF.R:35-67
  for (int t = 1; t <= T; t++) {
    const bool ident = (t == 1) || isnan(p.obs[(size_t)(t - 2) * p.dy]);     // R/CPF.R:38-40
    const double* y = p.obs + (size_t)(t - 1) * p.dy;
    const double sc = Model::step_const(c, t);
    // models whose weight factorises as exp(a) f (Model::SPLIT_W) park f in C next to a in S: both arrays are dead once
    // all ancestors are known, which the jump-list path guarantees
    const bool split_w = Model::SPLIT_W && jump_list;
    const bool shuffled = !LEAN && p.shuffle && !ident;
    double m = -INFINITY;                                  // running max of this thread's log-weights
    // The step runs once with parallel-scan sums.  If a sorted uniform came within rounding distance of a CDF knot (near != 0
    // somewhere in the CTA), it runs a second time with the reference's own operation order (exact): see "ancestors" below.
    bool exact = false;
    for (;;) {
    double sumw = 1.0;
    m = -INFINITY;
    if (tid == 0 && !exact) sm.iscr[35] = 0;             // near-knot flag (set by check_knots, read by vote_repeat); the repeat pass ignores it
    // block-wide vote at a barrier the step has anyway; true => repeat the step in exact order
    auto vote_repeat = [&]() -> bool { __syncthreads(); return sm.iscr[35] != 0 && !exact; };

    // -- sorted uniforms (src/multinomial.cpp:18-24) and CDF (:17); zero the offspring counters
    if (tree) for (int i = tid; i < (Npad >> 1); i += NT) sm.off32[i] = 0u;
    const bool draw_res = !ident && !inj_res;
    // Jump-list models (levy_sv): a particle's jump count comes from the second half of its resampling block.  The particles
    // that jump (31 % at the reference's theta) are appended to a queue of the WARP that owns them (4 words per entry in an
    // L2-resident scratch row), so that the jumps are later computed by full warps instead of the 16 of 32 lanes a per-thread
    // walk keeps busy.  hasj: bit e = particle e of this thread has jumps (its sums will be in its S / C slots).
    unsigned hasj = 0u;
    int nq1 = 0;                                           // entries in this warp's queue (warp-uniform)
    uint4* jq = Model::JUMP_LIST ? p.jq + (size_t)ws * pf_jq_stride(Npad) + (size_t)(tid >> 5) * 32 * ((N + NT - 1) / NT) : nullptr;   // a warp owns 32 x ceil(N / NT) particles
    if (draw_res && tid == 0) sm.dscr2[0] = -rng_log(philox_u1(stream, N, t, CPIMH_P_RESAMPLE, 0));   // E_N, read after the scan's barriers
    if constexpr (!Model::JUMP_LIST) {
      if (draw_res) for (int i = tid, ip = kp0; i < N; i += NT, ip += NTP) sm.S[ip] = -rng_log(philox_u1(stream, i, t, CPIMH_P_RESAMPLE, 0));   // exponential spacing E_i
    } else if (draw_res || (need_shared_u && jump_list)) {
      int e = 0;
      const int lane = tid & 31;
      for (int i0 = tid - lane; i0 < N; i0 += NT, e++) {   // warp-uniform trip count: the queue append votes with all 32 lanes
        const int i = i0 + lane, ip = kp0 + e * NTP;
        const bool in = i < N;
        double ua = 0.5; uint2 zw = make_uint2(0u, 0u);
        if (in) philox_u1raw(stream, i, t, CPIMH_P_RESAMPLE, 0, ua, zw);
        if (draw_res && in) sm.S[ip] = -rng_log(ua);                 // exponential spacing E_i (see the ladder comment below)
        if constexpr (Model::JUMP_LIST) {
          if (jump_list && need_shared_u) {               // the model turns its uniform into a small count right away
            int tk = in ? Model::token(c, zw) : 0;
            if (tk > 255) { tk = 255; flag_status(CPIMH_ERR_INVALID); }
            const unsigned vote = __ballot_sync(0xffffffffu, tk > 0);
            if (tk > 0) {
              hasj |= 1u << e;
              jq[nq1 + __popc(vote & ((1u << lane) - 1u))] = make_uint4((unsigned)i | ((unsigned)tk << 24), zw.x, zw.y, (unsigned)ip);
            }
            nq1 += __popc(vote);
          }
        }
      }
    }
    if (!ident) {
      if (inj_res) for (int i = tid; i < N; i += NT) sm.S[pad16(i)] = p.inj_resample[((size_t)fb * T + (t - 1)) * N + i];
      if (exact) for (int i = tid; i < N; i += NT) sm.C[pad16(i)] = p.wscr[(size_t)ws * Npad + i] / sm.dscr2[1];      // normweights, R/CPF.R:63 (dscr2[1]: last step's sum)
      __syncthreads();
      // Sorted uniforms.  Injected: S holds the ascending e_i of src/multinomial.cpp:22-24 and u_i = sumw e_i (:24).
      // Philox: the same order statistics from exponential spacings, e_i = (E_0+..+E_i) / (E_0+..+E_N) (one log per
      // particle instead of the recurrence's log + divide + exp; DESIGN.md "RNG"): S becomes the prefix sums P_i and
      // u_i = (sumw / total) P_i.
      if (!exact) {
        if (inj_res) block_scan_chunked<false>(sm.C, nullptr, N, sm.dscr);
        else block_scan_chunked<true>(sm.C, sm.S, N, sm.dscr);
      } else {                                             // Rcpp::cumsum's order (src/multinomial.cpp:17): one thread, left to right
        if (tid == 0) sequential_cumsum_p(sm.C, N);
        if (!inj_res && tid == (NT > 32 ? 32 : 0)) sequential_cumsum_p(sm.S, N);
        if (tid == 0) sm.iscr[34] += 1;                    // exact-order repeats of this filter
        __syncthreads();
      }
      sumw = sm.C[pad16(N - 1)];
      // A knot within `margin` of a uniform could fall on its other side under another summation order: two orders of the same
      // non-negative terms differ by at most 2 gamma_N of the total, likewise the two u_k; 12 (N + 2) 2^-53 sum(w) bounds the
      // sum of all those differences (DESIGN.md "Exact ancestors").
      if (tid == 0) sm.dscr2[2] = exact ? 0.0 : (double)(N + 2) * 0x1.8p-50 * sumw;      // read after the barrier below
      if (!inj_res) sumw = sumw / (sm.S[pad16(N - 1)] + sm.dscr2[0]);   // total = P_{N-1} + E_N
      // every 32nd sorted uniform gets a full search; the others then search only between their neighbours' results
      const int nrows = (N + 31) >> 5;
      for (int r = tid; r <= nrows; r += NT) {
        int a = N;
        if (r < nrows) a = count_le_p(sm.C, 0, N, sumw * sm.S[pad16(r << 5)]);
        sm.coarse[r] = a;
      }
      __syncthreads();
    } else if (tree) {
      __syncthreads();
    }

    // -- ancestors: count of CDF knots <= u (:25-28), clamped.  The CDF above is a parallel scan, the reference's a
    // left-to-right sum.  When no u_k comes within `margin` of the knots on either side of it, the ancestors found here ARE the
    // reference's (whatever the search did on the way: the two knots are re-read and compared).  Otherwise (a few times per
    // 10^4 steps at N = 4096) the CTA repeats the step with normalised weights and left-to-right sums by one thread.
    auto check_knots = [&](int a, double u) {
      if (CPIMH_EXP & 4) return;
      const double prev = sm.C[pad16(a > 0 ? a - 1 : 0)], cur = sm.C[pad16(a)], margin = sm.dscr2[2];      // C carries a +infinity sentinel at N
      if ((a > 0 && u - prev < margin) || (cur - u <= margin)) sm.iscr[35] = 1;
    };
    auto ancestor_of = [&](int k, int kp) -> int {
      if (ident) return k;
      const double u = sumw * sm.S[kp];
      int a = count_le_p(sm.C, sm.coarse[k >> 5], sm.coarse[(k >> 5) + 1], u);
      check_knots(a, u);
      if (a >= N) { a = N - 1; flag_status(512); }
      return a;
    };
    // -- gather (R/CPF.R:41) -> rtransition (:44) -> dmeasurement (:60) ; path bookkeeping
    auto finish_particle = [&](int k, int kp, int a, const double* x) {
#pragma unroll
      for (int q = 0; q < D; q++) xnew[(size_t)q * Npad + k] = x[q];
      double lw;
      if constexpr (Model::SPLIT_W) {
        if (split_w) {
          // a <= 0 by construction (Model::SPLIT_W): exp(a) cannot overflow, so the weight is formed right here with max = 0
          // and the separate weight pass is skipped; if EVERY a is below -600 (an observation no particle explains) the pass
          // below redoes the weights against the true maximum
          double f; Model::weight_parts(c, x, y, lw, f);
          const double w = exp(lw) * f;
          sm.C[kp] = w;
          if (!(CPIMH_EXP & 2)) p.wscr[(size_t)ws * Npad + k] = w;
        }
        else lw = Model::logw(c, x, y);
      } else lw = Model::logw(c, x, y);
      if (!split_w) sm.S[kp] = lw;
      m = fmax(m, lw);
      if (dense) hist_a[(size_t)(t - 1) * Npad + k] = a;
      if (tree) {                                      // src/tree.cpp:39-42, 90-94
        sm.bpar[k] = sm.leaf[a];
        atomicAdd(&sm.off32[a >> 1], 1u << ((a & 1) << 4));
      }
      if (!LEAN && p.anc_rec) p.anc_rec[((size_t)fb * T + (t - 1)) * N + k] = a;
    };
    auto propagate = [&](int k, int kp, int a) {
      double x[D];
#pragma unroll
      for (int q = 0; q < D; q++) x[q] = xold[(size_t)q * Npad + a];
      uint2 ush = make_uint2(0u, 0u);
      if constexpr (Model::USES_SHARED_U) {              // second half of the particle's OWN resampling block (DESIGN.md "RNG")
        if (!c.inj_trans) { double ua; philox_u1raw(stream, k, t, CPIMH_P_RESAMPLE, 0, ua, ush); }
      }
      flag_status(Model::transition(c, t, k, x, sc, ush));
      finish_particle(k, kp, a, x);
    };
    if (shuffled) {
      for (int k = tid, kp = kp0; k < N; k += NT, kp += NTP) sm.bpar[k] = ancestor_of(k, kp);
      if (vote_repeat()) { exact = true; continue; }
      // std::random_shuffle with randWrapper (src/multinomial.cpp:6-10,30): for i = 1..n-1 swap(a[i], a[floor(u_i (i+1))])
      int* jidx = reinterpret_cast<int*>(sm.S.ptr());
      for (int i = 1 + tid; i < N; i += NT) {
        double u = p.inj_shuffle ? p.inj_shuffle[((size_t)fb * T + (t - 1)) * (N - 1) + (i - 1)] : philox_u1(stream, i, t, CPIMH_P_SHUFFLE, 0);
        int j = (int)floor(u * (double)(i + 1));
        jidx[i] = j > i ? i : j;
      }
      // The N-1 swaps are a serial chain on the array; replayed in parallel instead by tracing every FINAL position back to
      // the original position its content came from.  Swap i touches positions i and j_i <= i, so going back in time from
      // position cur with swaps < m still to undo: if some swap i in (cur, m) has j_i == cur, the content sat at position i
      // before it and no earlier swap touches i: origin i (the largest such i); else swap cur itself moved it from j_cur:
      // continue from there with m = cur.  The swaps that hit a position are kept as buckets sorted descending (counting
      // sort + insertion sort of buckets of expected length <= ln N); ~1 hop per position for uniform draws.  Scratch: the
      // CDF and ladder arrays, both dead once the ancestors sit in bpar.  Buckets longer than 32 (only injected uniforms can
      // do that) and filters of fewer than 1024 particles use the one-thread loop.
      int* start = jidx + N;                               // [N + 1] bucket offsets (ladder array, second half)
      int* items = reinterpret_cast<int*>(sm.C.ptr());     // [N] swap indices grouped by the position they hit
      int* outp = items + N;                               // [N] fill cursors, then the permuted ancestors
      for (int q = tid; q <= N; q += NT) start[q] = 0;
      __syncthreads();
      for (int i = 1 + tid; i < N; i += NT) { const int j = jidx[i]; if (j != i) atomicAdd(&start[j], 1); }
      __syncthreads();
      const int cpt = (N + NT - 1) / NT;
      const int q0 = min(N, tid * cpt), q1 = min(N, q0 + cpt);
      int tot = 0, longest = 0;
      for (int q = q0; q < q1; q++) { const int cnt = start[q]; longest = max(longest, cnt); tot += cnt; }
      int total = 0;
      int off = block_excl_scan_int(tot, sm.iscr.ptr(), &total);
      if (__syncthreads_or(longest > 32 || N < 1024)) {     // small N: the 6 passes + barriers cost more than N serial swaps (gss N = 256: 1.4x)
        if (tid == 0)
          for (int i = 1; i < N; i++) { int j = jidx[i]; int tmp = sm.bpar[i]; sm.bpar[i] = sm.bpar[j]; sm.bpar[j] = tmp; }
      } else {
        for (int q = q0; q < q1; q++) { const int cnt = start[q]; start[q] = off; outp[q] = off; off += cnt; }
        if (tid == 0) start[N] = total;
        __syncthreads();
        for (int i = 1 + tid; i < N; i += NT) { const int j = jidx[i]; if (j != i) items[atomicAdd(&outp[j], 1)] = i; }
        __syncthreads();
        for (int q = tid; q < N; q += NT) {
          const int b0 = start[q], b1 = start[q + 1];
          for (int x = b0 + 1; x < b1; x++) {
            const int v = items[x];
            int y = x - 1;
            while (y >= b0 && items[y] < v) { items[y + 1] = items[y]; y--; }
            items[y + 1] = v;
          }
        }
        __syncthreads();
        for (int q = tid; q < N; q += NT) {
          int cur = q, m = N, origin;
          for (;;) {
            int found = -1;
            for (int x = start[cur], xe = start[cur + 1]; x < xe; x++) { const int it = items[x]; if (it < m) { found = it; break; } }
            if (found >= 0) { origin = found; break; }
            if (cur >= 1 && cur < m) { m = cur; cur = jidx[cur]; } else { origin = cur; break; }
          }
          outp[q] = sm.bpar[origin];
        }
        __syncthreads();
        for (int q = tid; q < N; q += NT) sm.bpar[q] = outp[q];
      }
      __syncthreads();
    }
    bool done = false;
    if constexpr (Model::JUMP_LIST) {
      if (jump_list) {
        // Models whose transition noise is a variable number of jumps per particle (Poisson counts: ~18 % of the lanes
        // would be active in a per-particle loop).  (1) all ancestors first, so the CDF and the ladder are dead;
        // (2) every thread walks the jumps of ALL its particles as one list, parking each particle's sums in that
        // particle's own slots of S and C; (3) gather, finish the transition, weigh.
        int* arow = dense ? hist_a + (size_t)(t - 1) * Npad : p.ascr + (size_t)ws * Npad;
        if (shuffled || ident) {
          for (int k = tid; k < N; k += NT) arow[k] = shuffled ? sm.bpar[k] : k;
        } else {
          // two of the thread's searches at a time (independent chains of shared-memory loads)
          for (int k = tid, kp = kp0; k < N; k += 2 * NT, kp += 2 * NTP) {
            const int k2 = k + NT;
            const bool two = k2 < N;
            const int kk = two ? k2 : k;
            int a1, a2;
            const double u1 = sumw * sm.S[kp], u2 = sumw * sm.S[two ? kp + NTP : kp];
            count_le_p2(sm.C, sm.coarse[k >> 5], sm.coarse[(k >> 5) + 1], u1, sm.coarse[kk >> 5], sm.coarse[(kk >> 5) + 1], u2, a1, a2);
            check_knots(a1, u1);
            check_knots(a2, u2);
            if (a1 >= N) { a1 = N - 1; flag_status(512); }
            if (a2 >= N) { a2 = N - 1; flag_status(512); }
            arow[k] = a1;
            if (two) arow[k2] = a2;
          }
        }
        if (vote_repeat()) { exact = true; continue; }
        const int E = (N - tid + NT - 1) / NT;
        if (!c.inj_trans) {
          // (1) first jumps: the warp's queue, one entry per lane and pass -- every lane busy, no Philox call (the entry carries the
          // two 32-bit uniforms of the particle's own resampling block); particles with more jumps go to a second queue
          const int lane = tid & 31;
          __syncwarp();
          int nq2 = 0;
          uint4 nxt = lane < nq1 ? jq[lane] : make_uint4(0u, 0u, 0u, 0u);
          for (int q0 = 0; q0 < nq1; q0 += 32) {
            const int q = q0 + lane;
            const bool on = q < nq1;
            const uint4 en = nxt;
            if (q + 32 < nq1) nxt = jq[q + 32];             // next pass's entry is in flight (L2) while this one is computed
            const int tk = (int)(en.x >> 24);
            if (on) {
              double wa, ev;
              Model::jump1(c, make_uint2(en.y, en.z), tk, wa, ev);
              sm.S[en.w] = wa; sm.C[en.w] = ev;
            }
            const unsigned more = __ballot_sync(0xffffffffu, on && tk > 1);
            __syncwarp();                                   // entry q0.. has been read by every lane before slot q0.. may be rewritten below
            if (on && tk > 1) jq[nq2 + __popc(more & ((1u << lane) - 1u))] = en;   // compacted in place: nq2 <= q0 + (lanes before me) <= q
            nq2 += __popc(more);
          }
          __syncwarp();
          // (2) jumps 2..k of the few particles that have them (6 % of the particles): one Philox block per jump, added in jump order
          for (int q0 = 0; q0 < nq2; q0 += 32) {
            const int q = q0 + lane;
            if (q < nq2) {
              const uint4 en = jq[q];
              const int tk = (int)(en.x >> 24), i = (int)(en.x & 0xffffffu);
              double sa = sm.S[en.w], sb = sm.C[en.w];
              for (int j = 2; j <= tk; j++) { double wa, ev; Model::jump(c, i, t, j, wa, ev); sa += wa; sb += ev; }
              sm.S[en.w] = sa; sm.C[en.w] = sb;
            }
          }
          __syncwarp();                                     // the sums of a particle were written by another lane of ITS OWN warp
        }
        // software pipeline over the thread's particles: the ancestor of particle e+2 and the gathered state of particle
        // e+1 are in flight (L2 latency) while particle e is finished and weighed
        int a_cur = E > 0 ? arow[tid] : 0, a_nxt = E > 1 ? arow[tid + NT] : 0;
        double x_nxt[D];
#pragma unroll
        for (int q = 0; q < D; q++) x_nxt[q] = E > 0 ? xold[(size_t)q * Npad + a_cur] : 0.0;
        for (int e = 0, kp = kp0; e < E; e++, kp += NTP) {
          const int k = tid + e * NT, a = a_cur;
          double x[D], swe = 0.0, se = 0.0;
#pragma unroll
          for (int q = 0; q < D; q++) x[q] = x_nxt[q];
          a_cur = a_nxt;
          if (e + 1 < E) {
#pragma unroll
            for (int q = 0; q < D; q++) x_nxt[q] = xold[(size_t)q * Npad + a_cur];
          }
          a_nxt = (e + 2 < E) ? arow[k + 2 * NT] : 0;
          if (c.inj_trans) { swe = trans_row(c, t, 0)[k]; se = trans_row(c, t, 1)[k]; }
          else if ((hasj >> e) & 1u) { swe = sm.S[kp]; se = sm.C[kp]; }
          Model::finish(c, x, swe, se);
          finish_particle(k, kp, a, x);
        }
        done = true;
      }
    }
    if (!done) {
      if constexpr (Model::MULTI) {
        if ((N + NT - 1) / NT <= CPIMH_MULTI_MAX_E) {
          // the thread's particles go through the transition together (see Model::transition_multi)
          double xl[CPIMH_MULTI_MAX_E][D];
          int al[CPIMH_MULTI_MAX_E];
          int E = 0;
          for (int k = tid, kp = kp0; k < N; k += NT, kp += NTP, E++) {
            const int a = shuffled ? sm.bpar[k] : ancestor_of(k, kp);
            al[E] = a;
#pragma unroll
            for (int q = 0; q < D; q++) xl[E][q] = xold[(size_t)q * Npad + a];
          }
          if (E > 0) flag_status(Model::transition_multi(c, t, E, tid, NT, xl, nullptr, sc));
          for (int e = 0; e < E; e++) finish_particle(tid + e * NT, kp0 + e * NTP, al[e], xl[e]);
          done = true;
        }
      }
    }
    if (!done) {
      for (int k = tid, kp = kp0; k < N; k += NT, kp += NTP) propagate(k, kp, shuffled ? sm.bpar[k] : ancestor_of(k, kp));
    }
    // the barrier that ends the propagation also carries the vote
    if (vote_repeat()) { exact = true; continue; }
    break;
    }

    // -- weights and log Z increment: R/CPF.R:61-63,66
    m = block_max(m, sm.dscr);
    double sloc = 0.0;
    double* wscr = p.wscr + (size_t)ws * Npad;             // this step's unnormalised weights, for an exact-order repeat of the next step
    if (split_w) {
      if constexpr (Model::SPLIT_W) {
        if (m >= -600.0) m = 0.0;                          // the weights are already in C and wscr (finish_particle)
        else {                                             // all weights would underflow: redo them relative to the maximum
          for (int k = tid, kp = kp0; k < N; k += NT, kp += NTP) {
            double x[D], a, f;
#pragma unroll
            for (int q = 0; q < D; q++) x[q] = xnew[(size_t)q * Npad + k];
            Model::weight_parts(c, x, y, a, f);
            const double w = exp(a - m) * f;
            sm.C[kp] = w; wscr[k] = w;
          }
        }
      }
    }
    else { for (int k = tid, kp = kp0; k < N; k += NT, kp += NTP) { double w = exp(sm.S[kp] - m); sm.C[kp] = w; if (!(CPIMH_EXP & 2)) wscr[k] = w; sloc += w; } }
    // the sum is evaluated as one fixed tree, whatever the CTA size (common.cuh block_sum_canonical)
    const double s = (CPIMH_EXP & 1) ? block_sum(sloc, sm.dscr) : block_sum_canonical([&](int, int kp) { return sm.C[kp]; }, N, sm.dscr);
    if (tid == 0) sm.dscr2[1] = s;
    // normweights = w / sum(w) (R/CPF.R:63).  Between steps only ratios matter (the ancestor search scales its
    // uniforms by the CDF total), so the division pass is done once, for the weights the filter returns.
    if (t == T) for (int k = tid; k < N; k += NT) sm.C[pad16(k)] = sm.C[pad16(k)] / s;
    if (tid == 0) {                                      // only thread 0 reports log Z
      double lz = log(s / (double)N) + m;
      if constexpr (Model::SPLIT_W) { if (split_w) lz += Model::w_const(); }
      sm.dscr2[3] += lz;
      if (p.logz_t) p.logz_t[(size_t)fb * T + (t - 1)] = lz;
    }

    // -- Tree::update (src/tree.cpp:88-99)
    if (tree) {
      // prune (:70-86): o_star <- scatter(o, l_star); walk up from childless leaves, decrementing offspring counts
      for (int i = tid; i < N; i += NT) {
        int o = (sm.off32[i >> 1] >> ((i & 1) << 4)) & 0xffff;
        int slot = sm.leaf[i];
        o_star[slot] = o;
        if (o == 0) {
          set_free_bit(sm.bits, slot);
          int j = a_star[slot];
          while (j >= 0) {
            int old = atomicSub(&o_star[j], 1);
            if (old != 1) break;
            set_free_bit(sm.bits, j);
            j = a_star[j];
          }
        }
      }
      __syncthreads();
      // insert (:36-68): first N free slots in index order (:45-53) from the shared-memory bitmap
      int found = 0;
      for (int base = 0; base < W && found < N; base += NT) {
        int wi = base + tid;
        unsigned w = wi < W ? sm.bits[wi] : 0u;
        int tot;
        int r = found + block_excl_scan_int(__popc(w), sm.iscr, &tot);
        while (w && r < N) {
          int b = __ffs(w) - 1;
          sm.leaf[r] = (wi << 5) + b;
          w &= w - 1;
          r++;
        }
        if (wi < W) sm.bits[wi] = w;
        found += tot;
      }
      if (found < N) {                      // Tree::double_size (:101-116) would be needed: report, host retries with 2M
        if (tid == 0) { *out.status = CPIMH_ERR_TREE_OVERFLOW; *out.logz = nan(""); }
        __syncthreads();
        return;
      }
      __syncthreads();
      for (int k = tid; k < N; k += NT) {   // a_star <- scatter(b, l_star); x_star <- scatter(x, l_star) (:64-67)
        int slot = sm.leaf[k];
        a_star[slot] = sm.bpar[k];
#pragma unroll
        for (int q = 0; q < D; q++) x_star[(size_t)q * M + slot] = xnew[(size_t)q * Npad + k];
      }
    }
    if (dense) { xold = xnew; xnew = xnew + xstride; }
    else { double* tmp = xold; xold = xnew; xnew = tmp; }
    __syncthreads();
  }

  // ---------------- path selection: systematic_resampling_n(normweights, 1, runif(1)) + Tree$get_path (R/CPF.R:69)
  if (p.normw) for (int k = tid; k < N; k += NT) p.normw[(size_t)fb * N + k] = sm.C[pad16(k)];
  if (p.leaf_out && tree) for (int k = tid; k < N; k += NT) p.leaf_out[(size_t)ws * N + k] = sm.leaf[k];
  if (p.exp_path && (tree || dense)) {
    // pf_get_weighted_average (R/CPF.R:84-92) over sapply(0:(N-1), Tree$get_path) (:73-74): all N leaves walk to the root
    // in lock step; per generation and component one block reduction of w_n x.  S is free and holds the walkers.
    int* cur = reinterpret_cast<int*>(sm.S.ptr());
    const double* hx = dense ? p.hist_x + (size_t)ws * (T + 1) * xstride : nullptr;
    for (int n = tid; n < N; n += NT) cur[n] = tree ? sm.leaf[n] : n;
    for (int step = T; step >= 0; step--) {
#pragma unroll
      for (int q = 0; q < D; q++) {
        const double acc = block_sum_canonical([&](int n, int np) {
          return (tree ? x_star[(size_t)q * M + cur[n]] : hx[(size_t)step * xstride + (size_t)q * Npad + cur[n]]) * sm.C[np]; }, N, sm.dscr);
        if (tid == 0) p.exp_path[((size_t)fb * (T + 1) + step) * D + q] = acc;
      }
      if (step > 0) for (int n = tid; n < N; n += NT) cur[n] = tree ? a_star[cur[n]] : hist_a[(size_t)(step - 1) * Npad + cur[n]];
    }
  }
  __syncthreads();
  if (tid == 0) sequential_cumsum_p(sm.C, N);   // csw of src/systematic.cpp:13-17: left to right, once per filter
  __syncthreads();
  double u = (!LEAN && p.inj_path) ? p.inj_path[fb] : philox_u1(stream, 0, T + 1, CPIMH_P_PATH, 0);
  int kidx = count_lt_p(sm.C, N, u);          // smallest j with csw_j >= u (src/systematic.cpp:15-18), n = 1
  int flags = sm.iscr[32];
  if (kidx >= N) { kidx = N - 1; flags |= 512; }
  if (tid == 0) {
    *out.logz = sm.dscr2[3];
    if (p.path_index) p.path_index[fb] = kidx;
    *out.status = flags & 255;
    if (p.exact_repeats) p.exact_repeats[fb] = sm.iscr[34];
  }
  if (out.traj && tid < 32) {               // Tree::get_path (src/tree.cpp:118-129): one warp, lane q carries component q
    if (tree) {
      int j = sm.leaf[kidx];
      for (int step = T; step >= 0; step--) {
        for (int q = tid; q < D; q += 32) out.traj[(size_t)step * D + q] = x_star[(size_t)q * M + j];
        j = a_star[j];
      }
    } else if (dense) {
      const double* hx = p.hist_x + (size_t)ws * (T + 1) * xstride;
      int j = kidx;
      for (int step = T; step >= 0; step--) {
        for (int q = tid; q < D; q += 32) out.traj[(size_t)step * D + q] = hx[(size_t)step * xstride + (size_t)q * Npad + j];
        if (step > 0) j = hist_a[(size_t)(step - 1) * Npad + j];
      }
    }
  }
  __syncthreads();
}

template <class Model, int NT_MAX, int MIN_BLOCKS, bool LEAN>
__global__ void __launch_bounds__(NT_MAX, MIN_BLOCKS) pf_batch_kernel(PfParams p) {
  PfSmem sm = pf_carve(p.lay);
  for (int i = threadIdx.x; i < CPIMH_MAX_THETA; i += blockDim.x) sm.theta[i] = p.theta[i];
  for (int i = threadIdx.x; i < p.npre; i += blockDim.x) sm.pre[i] = p.pre[i];
  __syncthreads();
  for (long long b = p.b0 + blockIdx.x; b < p.b0 + p.B; b += gridDim.x) {
    Stream s = make_stream(p.seed, p.stream_ids ? p.stream_ids[b] : p.first_filter + (unsigned long long)b);
    FilterOut out;
    out.logz = p.logz + b;
    out.traj = p.traj ? p.traj + (size_t)b * (p.T + 1) * Model::D : nullptr;
    out.status = p.status + b;
    run_filter<Model, LEAN>(p, sm, blockIdx.x, s, b, out);
  }
}

// host-side launcher shared by the per-model translation units.  Kernel variants by CTA size / residency target.
struct PfLaunch { int threads; int grid; size_t smem; int variant; };
enum { PF_VARIANT_128x8 = 0, PF_VARIANT_256x3 = 1, PF_VARIANT_512x2 = 2, PF_VARIANT_1024x1 = 3, PF_NUM_VARIANTS = 4 };
inline int pf_variant_threads(int v) { return v == PF_VARIANT_128x8 ? 128 : v == PF_VARIANT_256x3 ? 256 : v == PF_VARIANT_512x2 ? 512 : 1024; }
inline int pf_variant_for(int threads) { return threads <= 128 ? PF_VARIANT_128x8 : threads <= 256 ? PF_VARIANT_256x3 : threads <= 512 ? PF_VARIANT_512x2 : PF_VARIANT_1024x1; }

template <class Model, bool LEAN, class F>
int pf_with_kernel(int variant, F&& f) {
  switch (variant) {
    // wide states (d > 4) keep x, noise and the matvec in registers: trade residency for 128 registers per thread; the
    // narrow ones run best at ~80 registers (768 resident threads per SM): 6 x 128 or 3 x 256 (the enum names are historic)
    case PF_VARIANT_128x8: return f(pf_batch_kernel<Model, 128, (Model::D > 4 ? 5 : 6), LEAN>);
    case PF_VARIANT_256x3: return f(pf_batch_kernel<Model, 256, (Model::D > 4 ? 2 : 3), LEAN>);
    case PF_VARIANT_512x2: return f(pf_batch_kernel<Model, 512, (Model::D > 4 ? 1 : 2), LEAN>);
    default: return f(pf_batch_kernel<Model, 1024, 1, LEAN>);
  }
}
// the production configuration runs the lean instantiation (see run_filter)
inline bool pf_is_lean(const PfParams& p) {
  return p.path_mode == PATH_DENSE && !p.shuffle && !p.anc_rec && !p.inj_init && !p.inj_trans && !p.inj_resample && !p.inj_shuffle && !p.inj_path;
}
template <class Model>
int launch_pf_batch(const PfParams& p, const PfLaunch& L, cudaStream_t stream) {
  auto go = [&](auto kernel) -> int {
    CPIMH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    kernel<<<L.grid, L.threads, L.smem, stream>>>(p);
    CPIMH_CUDA_CHECK(cudaGetLastError());
    return CPIMH_OK;
  };
  return pf_is_lean(p) ? pf_with_kernel<Model, true>(L.variant, go) : pf_with_kernel<Model, false>(L.variant, go);
}
template <class Model>
int occupancy_pf_batch(int variant, int threads, size_t smem, int* blocks_per_sm) {
  return pf_with_kernel<Model, false>(variant, [&](auto kernel) -> int {
    CPIMH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPIMH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, threads, smem));
    return CPIMH_OK;
  });
}

}  // namespace cpimh
