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

// Number of SMs of the current device (cached per process).
int sm_count();

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
