// chains.cuh -- coupled PIMH chains + unbiased (H_bar) estimator as ONE persistent kernel: no host round trips.
//
// Restates, batched on the GPU, the reference's R-level drivers:
//   pimh_init / single_kernel / coupled_kernel   R/pf_kernels.R:86-145   (one CPF proposal, one uniform,
//                                                accept iff log(u) < logz_prop - logz_i, same u for both chains)
//   coupled_pimh_chains                          R/debiasing_algorithm.R:140-265 (lag-1 coupling, stop at max(tau, K))
//   H_bar(c_chains, h = identity, k, K)          R/debiasing_algorithm.R:284-316
// PIMH proposals do not depend on the chain state (R/pf_kernels.R:104-131 runs CPF() and then only compares log Z), so a
// replicate is: init filter, then one filter per iteration with an O(1) accept / meet recursion that has to see the
// proposals in order.  Scheduling, all on the device:
//   * OWNERS.  Every resident CTA takes replicates from an atomic ticket counter and owns one at a time: it runs the
//     replicate's proposals and replays the recursion.  Short (tau = 1) and long replicates balance across the SMs and no
//     wave of filters ever waits for another (the round-based driver of round 1 lost ~6 % to partial waves and host syncs).
//   * HELPERS.  A CTA that finds the ticket counter exhausted does not exit: it looks for the unfinished replicate with
//     the fewest proposals in flight, claims that replicate's next iteration (atomicCAS on the slot's counter), runs the
//     filter with its own HBM workspace and publishes the result in the owner's proposal ring.  The owner consumes ring
//     entries strictly in iteration order, so the chain and H_bar are bit-identical to a sequential run; proposals that
//     turn out to lie beyond the stopping time max(tau, K) are discarded (at most RING - 1 per replicate, run by CTAs that
//     had nothing else to do).  A replicate with meeting time 80 therefore does not serialise 80 filters on one CTA.
// The second filter that pimh_init runs and coupled_pimh_chains discards (R/pf_kernels.R:140-142 vs
// debiasing_algorithm.R:160-161) is NOT run; `nfilters` counts the filters the chains needed, `nfilters[1]` those actually run.
#pragma once
#include "pf_fused.cuh"

namespace cpimh {

#define CPIMH_RING 128         // proposals of one replicate that may be in flight / waiting for the in-order replay (a power of two)

struct ChainParams {
  PfParams pf;                 // workspace slot = blockIdx.x.  pf.traj / pf.logz / pf.status / pf.exp_path ARE the proposal rings:
                               //   entry slot * RING + (i mod RING) holds iteration i of the replicate that slot owns
  long long R;
  unsigned first_replicate;
  int k, K, max_iterations;    // max_iterations < 0: unbounded (R: Inf)
  double* state;               // [grid][2][T+1][D]   chain 1, chain 2
  double* acc;                 // [grid][2][T+1][D]   MCMC sum, bias-correction sum (all state components)
  double* estate;              // [grid][2][T+1][D]   all_particles = TRUE: the weighted-average paths that travel with the states; null when off
  double* eacc;                // [grid][2][T+1][D]
  int* ring_tag;               // [grid][RING]        iteration whose results an entry holds (-1 none)
  int* next_iter;              // [grid]              next iteration of the slot's replicate to hand out
  int* iter_done;              // [grid]              iterations the owner has replayed
  int* slot_rep;               // [grid]              replicate (0-based within the call) the slot owns, -1 none
  int* unfinished;             // [1]                 replicates not finished yet
  int* work_counter;           // [1]
  double* hbar;                // [R][T+1][D] (R layout d x (T+1) column-major per replicate)
  double* bc;                  // [R][T+1][D] or null: bias-correction term alone (BC(), R/debiasing_algorithm.R:330-363)
  double* hbar_rb;             // [R][T+1][D] or null: H_bar over the all-particle averages
  int* meetingtime;            // [R]
  int* iteration;              // [R]
  int* status;                 // [R]
  unsigned long long* nfilters;// [2]   filters the replicates needed (1 + iterations), filters actually run (incl. discarded speculative ones)
  double* samples1; double* samples2; double* log_pdf1; double* log_pdf2;   // recording (nullable)
  int max_rec;
};

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }

// claim the next iteration of the replicate in `slot` (thread 0 only): -1 when its ring is full or max_iterations is reached
__device__ __forceinline__ int chains_try_claim(const ChainParams& cp, int slot) {
  for (;;) {
    const int old = ld_volatile(cp.next_iter + slot), done = ld_volatile(cp.iter_done + slot);
    if (old - done >= CPIMH_RING) return -1;
    if (cp.max_iterations >= 0 && old > cp.max_iterations) return -1;
    if (atomicCAS(cp.next_iter + slot, old, old + 1) == old) return old;
  }
}

template <class Model, int NT_MAX, int MIN_BLOCKS, bool LEAN>
__global__ void __launch_bounds__(NT_MAX, MIN_BLOCKS) chains_kernel(ChainParams cp) {
  constexpr int D = Model::D;
  const PfParams& p = cp.pf;
  PfSmem sm = pf_carve(p.lay);
  const int tid = threadIdx.x, NT = blockDim.x, P = p.T + 1, PD = P * D;
  for (int i = tid; i < CPIMH_MAX_THETA; i += NT) sm.theta[i] = p.theta[i];
  for (int i = tid; i < p.npre; i += NT) sm.pre[i] = p.pre[i];
  __syncthreads();
  const int slot = blockIdx.x, G = gridDim.x;
  const long long ws = blockIdx.x;
  double* x1 = cp.state + (size_t)slot * 2 * PD;
  double* x2 = x1 + PD;
  double* accH = cp.acc + (size_t)slot * 2 * PD;
  double* accD = accH + PD;
  double* e1 = cp.estate ? cp.estate + (size_t)slot * 2 * PD : nullptr;
  double* e2 = e1 ? e1 + PD : nullptr;
  double* accHe = cp.eacc ? cp.eacc + (size_t)slot * 2 * PD : nullptr;
  double* accDe = accHe ? accHe + PD : nullptr;

  // The replicate's scalars live in shared memory (sm.own / ist), not in registers: nothing is held across run_filter.
  //   own[0] lz1, own[1] lz2;  ist[0] iter, [1] meet, [2] tau, [3] id1, [4] id2, [5] status, [6] finished, [7] a1, [8] a2,
  //   [9] mode (0 needs a replicate, 1 owns one, 2 helper), [10] replicate owned, [11..13] work item (slot, iteration, replicate)
  int* ist = reinterpret_cast<int*>(&sm.own[4]);
  if (tid == 0) { ist[9] = 0; ist[10] = -1; }
  __syncthreads();

  // ONE loop with ONE run_filter call site (the filter is ~100 KB of code: a second inlined copy would thrash the instruction
  // cache): every pass picks a work item -- as owner: the next replicate's init filter or the owned replicate's next proposal;
  // as helper: a proposal of somebody else's replicate -- runs it, and, as owner, replays whatever has arrived.
  for (;;) {
    // ---------------------------------------------------------------- pick the work item
    if (ist[9] == 0) {                                                            // owner without a replicate: take a ticket
      __syncthreads();
      if (tid == 0) {
        const long long r = atomicAdd(cp.work_counter, 1);
        if (r >= cp.R) { ist[9] = 2; ist[10] = -1; }
        else { ist[9] = 1; ist[10] = (int)r; ist[11] = slot; ist[12] = 0; ist[13] = (int)r; cp.next_iter[slot] = 1; cp.iter_done[slot] = 0; }
      }
      // nobody else touches this slot now: helpers exist only once the tickets have run out, and then no slot starts a new replicate
      for (int i = tid; i < CPIMH_RING; i += NT) cp.ring_tag[slot * CPIMH_RING + i] = -1;
      __syncthreads();
      // publish the replicate BEFORE its init filter runs: proposals do not depend on the chain state, so helpers may already
      // run iterations 1, 2, ... while the owner runs iteration 0 (shortens the tail of a call by one filter time)
      if (tid == 0 && ist[9] == 1) { __threadfence(); st_volatile(cp.slot_rep + slot, ist[10]); }
    } else if (ist[9] == 1) {                                                     // owner: claim the replicate's next proposal
      __syncthreads();
      if (tid == 0) { ist[11] = slot; ist[12] = chains_try_claim(cp, slot); ist[13] = ist[10]; }
      __syncthreads();
    }
    if (ist[9] == 2) {                                                            // helper: the unfinished replicate with the fewest proposals in flight
      int best = 0x7fffffff;                                                      // key = lookahead * G + slot
      for (int s = tid; s < G; s += NT) {
        if (ld_volatile(cp.slot_rep + s) < 0) continue;
        const int nxt = ld_volatile(cp.next_iter + s), la = nxt - ld_volatile(cp.iter_done + s);
        if (la >= CPIMH_RING || (cp.max_iterations >= 0 && nxt > cp.max_iterations)) continue;
        best = min(best, la * G + s);
      }
      for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      __syncthreads();
      if ((tid & 31) == 0) sm.iscr[tid >> 5] = best;
      __syncthreads();
      if (tid == 0) {
        int b = 0x7fffffff;
        for (int w = 0; w < (NT >> 5); w++) b = min(b, sm.iscr[w]);
        int s = -1, it = -1, rp = -1;
        if (b != 0x7fffffff) {
          s = b % G;
          rp = ld_volatile(cp.slot_rep + s);
          it = rp >= 0 ? chains_try_claim(cp, s) : -1;
        }
        ist[11] = s; ist[12] = it > 0 ? it : -1; ist[13] = rp;
        if (it <= 0 && ld_volatile(cp.unfinished) <= 0) ist[12] = -2;             // everything is finished: leave
      }
      __syncthreads();
      if (ist[12] == -2) break;
      if (ist[12] < 0) { __nanosleep(2000); continue; }
    }
    // ---------------------------------------------------------------- run it: results into ring entry (s, it mod RING), then the tag
    if (ist[12] >= 0) {
      const int s = ist[11], it = ist[12];
      const unsigned rep = cp.first_replicate + (unsigned)ist[13];
      const long long e = (long long)s * CPIMH_RING + (it & (CPIMH_RING - 1));
      FilterOut out;
      out.logz = p.logz + e; out.traj = p.traj + (size_t)e * PD; out.status = p.status + e;
      run_filter<Model, LEAN>(p, sm, ws, make_stream(p.seed, (unsigned long long)rep | ((unsigned long long)it << 32)), e, out);   // ends with a barrier
      if (tid == 0) {
        atomicAdd(cp.nfilters + 1, 1ull);                  // filters actually run, speculative ones included
        __threadfence();                                   // results before the tag (the barrier made the other threads' writes ours)
        st_volatile(cp.ring_tag + e, it);
      }
    }
    if (ist[9] != 1) continue;
    // ---------------------------------------------------------------- owner: replay, strictly in iteration order, what has arrived
    const long long r = ist[10];
    const unsigned rep = cp.first_replicate + (unsigned)r;
    if (ist[12] == 0) {                                                           // pimh_init (R/pf_kernels.R:134-139): iteration 0 becomes chain 1
      const long long e0 = (long long)slot * CPIMH_RING;
      for (int j = tid; j < PD; j += NT) {
        const double v = __ldcg(p.traj + (size_t)e0 * PD + j);
        x1[j] = v; accH[j] = (cp.k <= 0) ? v : 0.0; accD[j] = 0.0;
        if (e1) { const double ev = __ldcg(p.exp_path + (size_t)e0 * PD + j); e1[j] = ev; accHe[j] = (cp.k <= 0) ? ev : 0.0; accDe[j] = 0.0; }   // debiasing_algorithm.R:166-170
      }
      if (cp.samples1) for (int j = tid; j < P; j += NT) cp.samples1[((size_t)r * (cp.max_rec + 1)) * P + j] = __ldcg(p.traj + (size_t)e0 * PD + (size_t)j * D);
      if (tid == 0) {
        int status = __ldcg(p.status + e0);
        const double lz = __ldcg(p.logz + e0);
        if (status == 0 && isnan(lz)) status = CPIMH_ERR_INVALID;
        sm.own[0] = lz; sm.own[1] = -INFINITY;
        ist[0] = 0; ist[1] = 0; ist[2] = -1; ist[3] = 0; ist[4] = -1; ist[5] = status; ist[6] = 0;
        if (cp.samples1) cp.log_pdf1[(size_t)r * (cp.max_rec + 1)] = lz;
      }
      __syncthreads();
    }
    bool progressed = false;
    while (!ist[6] && !ist[5]) {                                                  // R/debiasing_algorithm.R:180
      const int iter = ist[0] + 1;
      const long long e = (long long)slot * CPIMH_RING + (iter & (CPIMH_RING - 1));
      __syncthreads();
      if (tid == 0) {
        const int ready = (ld_volatile(cp.ring_tag + e) == iter);
        if (ready) {
          __threadfence();
          int status = __ldcg(p.status + e);
          const double lzp = __ldcg(p.logz + e);
          if (status == 0 && isnan(lzp)) status = CPIMH_ERR_INVALID;              // all weights zero / NaN: no valid proposal
          ist[0] = iter; ist[5] = status;
          if (status == 0) {
            double lz1 = sm.own[0], lz2 = sm.own[1];
            int meet = ist[1], tau = ist[2], id1 = ist[3], id2 = ist[4];
            const Stream st = make_stream(p.seed, (unsigned long long)rep | ((unsigned long long)iter << 32));
            const double lu = log(philox_u1(st, 0, 0, CPIMH_P_MH, 0));            // R/pf_kernels.R:112
            const bool a1 = lu < (lzp - lz1);                                     // :113
            const bool a2 = meet ? a1 : (lu < (lzp - lz2));                       // :121 (lz2 = -Inf at iteration 1: always accepts)
            if (a1) { lz1 = lzp; id1 = iter; }
            if (meet) { lz2 = lz1; id2 = id1; }                                   // single_kernel: both chains move together
            else {
              if (a2) { lz2 = lzp; id2 = iter; }
              if (id1 == id2) { meet = 1; tau = iter; }                           // debiasing_algorithm.R:216-220
            }
            sm.own[0] = lz1; sm.own[1] = lz2;
            ist[1] = meet; ist[2] = tau; ist[3] = id1; ist[4] = id2; ist[7] = a1; ist[8] = a2;
            if (tau > 0 && iter >= max(tau, cp.K)) ist[6] = 1;                    // :245
            else if (cp.max_iterations >= 0 && iter >= cp.max_iterations) { ist[6] = 1; ist[5] = CPIMH_ERR_NOT_MET; }   // R: finished = FALSE
            if (cp.samples1 && iter <= cp.max_rec) { cp.log_pdf1[(size_t)r * (cp.max_rec + 1) + iter] = lz1; cp.log_pdf2[(size_t)r * cp.max_rec + (iter - 1)] = lz2; }
          }
        }
        sm.iscr[33] = ready;
      }
      __syncthreads();
      if (!sm.iscr[33]) break;
      progressed = true;
      if (ist[5] != 0 && ist[5] != CPIMH_ERR_NOT_MET) break;                      // the proposal failed
      {
        const bool a1 = ist[7] != 0, a2 = ist[8] != 0;
        // H_bar on the fly (debiasing_algorithm.R:296-315): MCMC term for k <= iter <= K, bias term t = iter-1 >= k with
        // coefficient min(t-k+1, K-k+1) (zero once the chains have met)
        const int t = iter - 1;
        const bool mcmc = (iter >= cp.k && iter <= cp.K);
        const bool bias = (t >= cp.k) && (ist[3] != ist[4]);
        const double coef = (double)min(t - cp.k + 1, cp.K - cp.k + 1);
        const double* prop = p.traj + (size_t)e * PD;
        for (int j = tid; j < PD; j += NT) {
          const double pv = __ldcg(prop + j);
          const double xa = a1 ? pv : x1[j];
          const double xb = a2 ? pv : x2[j];
          if (a1) x1[j] = pv;
          if (a2) x2[j] = pv;
          if (mcmc) accH[j] += xa;                                                // :296-302
          if (bias) accD[j] = accD[j] + coef * (xa - xb);                         // :308-312
        }
        if (e1) {                                                                 // exp_path follows the same accept decisions (pf_kernels.R:176-190)
          const double* pe = p.exp_path + (size_t)e * PD;
          for (int j = tid; j < PD; j += NT) {
            const double pv = __ldcg(pe + j);
            const double xa = a1 ? pv : e1[j];
            const double xb = a2 ? pv : e2[j];
            if (a1) e1[j] = pv;
            if (a2) e2[j] = pv;
            if (mcmc) accHe[j] += xa;
            if (bias) accDe[j] = accDe[j] + coef * (xa - xb);
          }
        }
        if (cp.samples1 && iter <= cp.max_rec) {                                  // :235-236 chain_state[1,]: first component only
          __syncthreads();                                                        // the states were updated with another striping
          for (int j = tid; j < P; j += NT) {
            cp.samples1[((size_t)r * (cp.max_rec + 1) + iter) * P + j] = x1[(size_t)j * D];
            cp.samples2[((size_t)r * cp.max_rec + (iter - 1)) * P + j] = x2[(size_t)j * D];
          }
        }
      }
      __syncthreads();                                                            // everyone is done with the ring entry
      if (tid == 0) { __threadfence(); st_volatile(cp.iter_done + slot, iter); }  // ... which frees it for iteration iter + RING
    }
    __syncthreads();
    if (ist[6] || ist[5]) {                                                       // the replicate is finished (or failed): results, then the next ticket
      if (tid == 0) st_volatile(cp.slot_rep + slot, -1);                          // no further claims
      const double denom = (double)(cp.K - cp.k + 1);
      for (int j = tid; j < PD; j += NT) {
        cp.hbar[(size_t)r * PD + j] = (accH[j] + accD[j]) / denom;                // :315
        if (cp.bc) cp.bc[(size_t)r * PD + j] = accD[j] / denom;
        if (cp.hbar_rb && e1) cp.hbar_rb[(size_t)r * PD + j] = (accHe[j] + accDe[j]) / denom;
      }
      __syncthreads();
      if (tid == 0) {
        cp.meetingtime[r] = ist[2];
        cp.iteration[r] = ist[0];
        cp.status[r] = ist[5];
        atomicAdd(cp.nfilters, (unsigned long long)(1 + ist[0]));                 // filters the replicate needed (the reference's count)
        __threadfence();
        atomicSub(cp.unfinished, 1);
        ist[9] = 0;
      }
      __syncthreads();
    } else if (ist[12] < 0 && !progressed) __nanosleep(500);                      // ring full of proposals still in flight
  }
}

template <class Model, bool LEAN, class F>
int chains_with_kernel(int variant, F&& f) {
  switch (variant) {
    case PF_VARIANT_128x8: return f(chains_kernel<Model, 128, (Model::D > 4 ? 5 : 6), LEAN>);
    case PF_VARIANT_256x3: return f(chains_kernel<Model, 256, (Model::D > 4 ? 2 : 3), LEAN>);
    case PF_VARIANT_512x2: return f(chains_kernel<Model, 512, (Model::D > 4 ? 1 : 2), LEAN>);
    default: return f(chains_kernel<Model, 1024, 1, LEAN>);
  }
}
template <class Model>
int launch_chains(const ChainParams& cp, const PfLaunch& L, cudaStream_t stream) {
  auto go = [&](auto kernel) -> int {
    CPIMH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem));
    kernel<<<L.grid, L.threads, L.smem, stream>>>(cp);
    CPIMH_CUDA_CHECK(cudaGetLastError());
    return CPIMH_OK;
  };
  return pf_is_lean(cp.pf) ? chains_with_kernel<Model, true>(L.variant, go) : chains_with_kernel<Model, false>(L.variant, go);
}
template <class Model>
int occupancy_chains(int variant, int threads, size_t smem, int* blocks_per_sm) {
  return chains_with_kernel<Model, false>(variant, [&](auto kernel) -> int {
    CPIMH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CPIMH_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, threads, smem));
    return CPIMH_OK;
  });
}

// ------------------------------------------------------------------ single model operations (primitive API)
struct ModelOpParams {
  int op;                      // 0 rinit, 1 rtransition, 2 dmeasurement
  int N, T, time;
  const double* theta; const double* pre; int npre;
  unsigned long long seed, filter_id;
  const double* noise;         // [n][N] component-major (device) or null
  int n_trans, sk_M, integrator, substeps;
  double* x;                   // d x N column-major (R layout), device
  const double* y;             // [dim_y]
  double* logw;                // [N]
  int* status;                 // [1]
};
template <class Model>
__global__ void model_op_kernel(ModelOpParams q) {
  constexpr int D = Model::D;
  ModelCtx c;                                  // dynamic window: theta [CPIMH_MAX_THETA] then pre [npre]
  c.theta.off = 0; c.pre.off = CPIMH_MAX_THETA * (int)sizeof(double);
  for (int i = threadIdx.x; i < CPIMH_MAX_THETA; i += blockDim.x) c.theta[i] = q.theta[i];
  for (int i = threadIdx.x; i < q.npre; i += blockDim.x) c.pre[i] = q.pre[i];
  __syncthreads();
  c.stream = make_stream(q.seed, q.filter_id); c.N = q.N; c.T = q.T;
  c.inj_init = q.op == 0 ? q.noise : nullptr;
  // rtransition with injected noise: present the single step as row `time` of a [T][n][N] array
  c.inj_trans = (q.op == 1 && q.noise) ? q.noise - (size_t)(q.time - 1) * q.n_trans * q.N : nullptr;
  c.n_trans = q.n_trans; c.sk_M = q.sk_M; c.integrator = q.integrator; c.substeps = q.substeps;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.N) return;
  double x[D];
  if (q.op != 0) {
#pragma unroll
    for (int k = 0; k < D; k++) x[k] = q.x[(size_t)i * D + k];
  }
  if (q.op == 0) Model::init(c, i, x);
  else if (q.op == 1) {
    double ua = 0.5; uint2 zw = make_uint2(0u, 0u);
    if (Model::USES_SHARED_U && !c.inj_trans) philox_u1raw(c.stream, i, q.time, CPIMH_P_RESAMPLE, 0, ua, zw);
    int st = Model::transition(c, q.time, i, x, Model::step_const(c, q.time), zw);
    if (st & 255) atomicOr(q.status, st & 255);
  }
  else { q.logw[i] = Model::logw(c, x, q.y); return; }
#pragma unroll
  for (int k = 0; k < D; k++) q.x[(size_t)i * D + k] = x[k];
}
template <class Model>
int launch_model_op(const ModelOpParams& q, cudaStream_t stream) {
  const int threads = 128, grid = (q.N + threads - 1) / threads;
  model_op_kernel<Model><<<grid, threads, (size_t)(CPIMH_MAX_THETA + q.npre + 1) * sizeof(double), stream>>>(q);
  CPIMH_CUDA_CHECK(cudaGetLastError());
  return CPIMH_OK;
}

}  // namespace cpimh
