// registry.h -- per-model kernel entry points (one translation unit per model keeps nvcc compile times parallel)
#pragma once
#include "pf_fused.cuh"
#include "chains.cuh"
#include "ccpf.cuh"
#include "apf.cuh"

namespace cpimh {

struct ModelEntry {
  int model;
  int dim;
  int (*launch_pf)(const PfParams&, const PfLaunch&, cudaStream_t);
  int (*occupancy_pf)(int variant, int threads, size_t smem, int* blocks_per_sm);
  int (*launch_chains)(const ChainParams&, const PfLaunch&, cudaStream_t);
  int (*occupancy_chains)(int variant, int threads, size_t smem, int* blocks_per_sm);
  int (*launch_model_op)(const ModelOpParams&, cudaStream_t);
  int (*launch_ccpf)(const CcpfParams&, const CcpfLaunch&, cudaStream_t);
  int (*occupancy_ccpf)(int threads, size_t smem, int* blocks_per_sm);
  int (*launch_ccpf_chains)(const CcpfParams&, const CcRound&, long long, int*, const CcpfLaunch&, cudaStream_t);
  int (*launch_apf)(const ApfParams&, const ApfLaunch&, cudaStream_t);
  int (*occupancy_apf)(int threads, size_t smem, int* blocks_per_sm);
};
template <class Model>
ModelEntry make_entry(int model) {
  ModelEntry e;
  e.model = model; e.dim = Model::D;
  e.launch_pf = &launch_pf_batch<Model>;
  e.occupancy_pf = &occupancy_pf_batch<Model>;
  e.launch_chains = &launch_chains<Model>;
  e.occupancy_chains = &occupancy_chains<Model>;
  e.launch_model_op = &launch_model_op<Model>;
  e.launch_ccpf = &launch_ccpf<Model>;
  e.occupancy_ccpf = &occupancy_ccpf<Model>;
  e.launch_ccpf_chains = &launch_ccpf_chains<Model>;
  e.launch_apf = &launch_apf<Model>;
  e.occupancy_apf = &occupancy_apf<Model>;
  return e;
}
const ModelEntry* entries_gss(int* n);
const ModelEntry* entries_ar1(int* n);
const ModelEntry* entries_ar2(int* n);
const ModelEntry* entries_ar3(int* n);
const ModelEntry* entries_ar4(int* n);
const ModelEntry* entries_ar5(int* n);
const ModelEntry* entries_ar6(int* n);
const ModelEntry* entries_ar7(int* n);
const ModelEntry* entries_ar8(int* n);
const ModelEntry* entries_ar9(int* n);
const ModelEntry* entries_ar10(int* n);
const ModelEntry* entries_ar11(int* n);
const ModelEntry* entries_ar12(int* n);
const ModelEntry* entries_ar13(int* n);
const ModelEntry* entries_ar14(int* n);
const ModelEntry* entries_ar15(int* n);
const ModelEntry* entries_ar16(int* n);
const ModelEntry* entries_sv(int* n);
const ModelEntry* entries_sk(int* n);
const ModelEntry* entries_lorenz4(int* n);
const ModelEntry* entries_lorenz8(int* n);
const ModelEntry* entries_lorenz16(int* n);
const ModelEntry* find_entry(int model, int dim);

}  // namespace cpimh
