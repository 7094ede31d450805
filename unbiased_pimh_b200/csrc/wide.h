// wide.h -- host interface of the wide-state filter (wide.cu), used by the C ABI in api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cpimh.h"

namespace cpimh {
struct WideFilter;
int wide_supported(const cpimh_config& c);
int wide_create(const cpimh_config& cfg, int64_t max_filters, WideFilter** out);
void wide_destroy(WideFilter* h);
int wide_set_observations(WideFilter* h, const double* obs_host);
int wide_set_noise(WideFilter* h, int64_t B, const cpimh_noise* nz);
int wide_record_ancestors(WideFilter* h, int on);
int wide_run(WideFilter* h, int64_t B, uint64_t seed, uint64_t first_filter, cudaStream_t s);
int wide_fetch(WideFilter* h, int64_t B, cudaStream_t s, double* logz, double* logz_t, double* trajectory, int32_t* path_index);
int wide_fetch_ancestors(WideFilter* h, int64_t b, cudaStream_t s, int32_t* anc);
void wide_exact_mode(WideFilter* h, int on);
int wide_all_particles(WideFilter* h, int on);
int wide_fetch_exp_paths(WideFilter* h, int64_t B, cudaStream_t s, double* exp_path);
double* wide_device_exp_paths(WideFilter* h);
int wide_fetch_exact_repeats(WideFilter* h, int64_t B, cudaStream_t s, int32_t* repeats);
int64_t wide_launch_count(const WideFilter* h);
int64_t wide_device_bytes(const WideFilter* h);
void wide_device_outputs(WideFilter* h, double** d_logz, double** d_traj);
void wide_effective(const WideFilter* h, cpimh_config* out);
}  // namespace cpimh
