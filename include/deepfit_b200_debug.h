/*
 * deepfit_b200 -- test and profiling hooks of libdeepfit_b200.so.  NOT part of the drop-in boundary
 * (include/deepfit_b200.h): nothing under deepfit_b200/ but the test-only bindings in _lib.py uses them, a host
 * that binds the ABI never needs them, and they may change between builds without an ABI version bump.
 *
 * They are process-global switches and read-outs (not per handle, not thread-safe) that exist so that the tests can
 * compare the fused kernels with their unfused building blocks on the same inputs, and so that tools/timeline.py can
 * read the per-CTA clock stamps.  bench.py refuses to report a line as valid when DFB_DBG (its way of flipping them)
 * is set: config.dfb_dbg records it.
 */
#ifndef DEEPFIT_B200_DEBUG_H_
#define DEEPFIT_B200_DEBUG_H_

#include "deepfit_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A/B switches of the stage-3 kernels (csrc/mlp.cu).  key: 0 swap LBO/SBO (descriptor probe), 1 resident max-pool
 * kernel (0 = one CTA per (patch, channel tile)), 2 fused head (0 = separate 64->512 / 512->256 launches), 4 layer 3
 * fused into the head, 5 fused chain kernels (0 = pointwise / small-GEMM / max-pool launches), 6 head experiment bits
 * (results are garbage), 7 CTA whose timeline is recorded (-1 off), 8 persistent head, 9 per-tile entry/exit trace of
 * the head, 10 persistent chain, 11 corrections-first product order in the generic GEMM kernel (0 = hh, hl, lh per
 * k-step in one pass over K).  Every default is the product path. */
DFB_API int dfb_dbg_set(int key, int value);

/* (SM id, entry ns, exit ns) of the first n_cta CTAs of the last traced head launch (key 9).  Synchronises the device. */
DFB_API int dfb_dbg_cta_trace(long long* host_out, int n_cta);

/* The 4 x 64 clock64() stamps (chains 0-2, head) of the traced CTA (key 7); n >= 256.  Synchronises the device. */
DFB_API int dfb_dbg_timeline(long long* host_out, int n);

/* Unpacks one activation buffer of the model's last forward to fp32 row-major d_out[rows, cols].
 * which: 0 x64a (pf conv1), 1 x64b (pf conv2), 2 x64c (x*T2), 3 x128 (enc conv2), 4 h512, 5 h256, 6 gfeat, 7 f512,
 * 8 f256; only buffers the current configuration materialises (the fused kernels keep most of them on chip). */
DFB_API int dfb_dbg_model_dump(dfb_model* model, int32_t which, int64_t rows, int32_t cols, float* d_out,
                               dfb_stream stream);

/* C[M,N] = A[M,K] B[N,K]^T (+ bias, ReLU) through the generic tcgen05 kernel, fp32 row-major in and out.
 * mode 0: N orientation, activation epilogue; mode 1: T orientation + max over groups of k_pts rows of A.
 * nprod: 1 single BF16 product, 2 f16f8, 3 bf16x3, 4 f16x3. */
DFB_API int dfb_dbg_gemm(const float* d_A, const float* d_B, const float* d_bias, int64_t M, int32_t N, int32_t K,
                         int32_t mode, int32_t nprod, int32_t relu, int32_t k_pts, float* d_out,
                         float* d_out_tiled_unpacked, dfb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPFIT_B200_DEBUG_H_ */
