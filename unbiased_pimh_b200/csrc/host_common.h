// host_common.h -- host-side helpers shared by the C-ABI translation units
#pragma once
#include <vector>
#include "registry.h"

namespace cpimh {

// cudaGetDeviceProperties costs milliseconds (it queries every attribute); the one-shot entries ask once per device
inline cudaError_t cached_device_properties(cudaDeviceProp* out, int dev) {
  static cudaDeviceProp props[64];
  static bool have[64] = {false};
  if (dev < 0 || dev >= 64) return cudaGetDeviceProperties(out, dev);
  if (!have[dev]) { cudaError_t e = cudaGetDeviceProperties(&props[dev], dev); if (e != cudaSuccess) return e; have[dev] = true; }
  *out = props[dev];
  return cudaSuccess;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n, int64_t* total = nullptr) {
    if (n <= bytes && p) return CPIMH_OK;
    if (p) { cudaFree(p); if (total) *total -= (int64_t)bytes; p = nullptr; bytes = 0; }
    if (n == 0) return CPIMH_OK;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaErrorMemoryAllocation) {        // drop the workspaces cached by the one-shot entries (api.cu) and try once more
      cudaGetLastError();
      cpimh_release_cached();
      e = cudaMalloc(&p, n);
    }
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu bytes) failed: %s", n, cudaGetErrorString(e));
      p = nullptr;
      return e == cudaErrorMemoryAllocation ? CPIMH_ERR_NOMEM : CPIMH_ERR_CUDA;
    }
    bytes = n;
    if (total) *total += (int64_t)n;
    return CPIMH_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// launch geometry shared by the filter and the chain kernels
struct Geometry {
  int Npad, M, threads, blocks_per_sm, num_sms, variant, path_mode;
  size_t smem;
  size_t hist_x_elems, hist_a_elems;   // per workspace (PATH_DENSE)
};

int api_validate(cpimh_config* c);
int api_plan_geometry(const cpimh_config& c, const ModelEntry* e, bool chains, int64_t max_items, int npre, Geometry* g);
int api_plan_tail_geometry(const cpimh_config& c, const ModelEntry* e, int npre, const Geometry& main, Geometry* tail);
void api_fill_pf_params(const cpimh_config& c, const Geometry& g, int npre, PfParams* p);
std::vector<double> api_precompute(const cpimh_config& c);
int api_upload_obs(const cpimh_config& cfg, const double* obs_host, void* dst);
int api_n_init(const cpimh_config& c);
int api_n_trans(const cpimh_config& c);
struct AllParticlesArgs {
  int N, D, M, T, Npad, path_mode;
  const int* a_star; const double* x_star; const int* leaf;      // PATH_TREE
  const double* hist_x; const int* hist_a;                       // PATH_DENSE
  const double* normw;
};
int api_all_particles(const AllParticlesArgs& a, cudaStream_t s, double* h_exp_path, double* h_paths, double* h_normw);

}  // namespace cpimh
