// Shared helpers for the deepfit_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deepfit_b200.h"

namespace dfb {

// Thread-local last-error text, exposed through dfb_last_error_string().
void set_error(const char* fmt, ...);

inline cudaStream_t as_stream(dfb_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// Returns 0 or a positive CUDA error code after recording the message.
int check_cuda(cudaError_t e, const char* what, const char* file, int line);

#define DFB_CUDA(expr)                                                        \
    do {                                                                      \
        int _rc = ::dfb::check_cuda((expr), #expr, __FILE__, __LINE__);       \
        if (_rc) return _rc;                                                  \
    } while (0)

#define DFB_REQUIRE(cond, code, ...)                                          \
    do {                                                                      \
        if (!(cond)) {                                                        \
            ::dfb::set_error(__VA_ARGS__);                                    \
            return (code);                                                    \
        }                                                                     \
    } while (0)

// Every kernel launch of the library is counted (dfb_launch_count) so callers can report how many
// of OUR kernels ran in a timed region.
void note_launch(int n = 1);

// Number of SMs of the current device (cached per device).
int sm_count();

// Raise a kernel's dynamic shared-memory limit on the CURRENT device (once per function and device).
int ensure_dyn_smem(const void* func, int bytes);

// Internal cross-file entry points (the C pipeline drives the stages through these and the public ABI).
int wls_fit_impl(const float* d_patch, const float* d_weights, const float* d_trans, const float* d_U, const double* d_rad,
                 int64_t n, int32_t k, int32_t order, float* d_beta, float* d_n_local, float* d_n_world, float* d_curv,
                 double* d_curv_scaled, float* d_dirs, float* d_dirs_world, int32_t* d_info, int keep_qstn, dfb_stream stream,
                 const int32_t* d_out_row = nullptr, int64_t out_base = 0);
const float* grid_points(const dfb_grid* g);   // device copy of the cloud, original order, f32[n,3]
const float4* grid_points4(const dfb_grid* g); // the same values as float4 (x, y, z, 0): one 16-byte load per gathered point
int patch_pca_impl(const float* d_points, const float4* d_points4, int64_t n_points, const int32_t* d_idx, const int32_t* d_valid,
                   const int32_t* d_query, int64_t q_begin, int64_t q_count, int32_t k, const double* d_rad, int32_t scale_by_rad,
                   float* d_patch, float* d_U, double* d_cov, dfb_stream stream);
long long grid_size(const dfb_grid* g);
int model_k(const dfb_model* m);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dfb
