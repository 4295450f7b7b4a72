// Stage 3: the point-weight network on the 5th-generation tensor cores (tcgen05 + TMEM + TMA bulk
// copies), sm_100a only.
//
// Replaces DeepFit.forward up to the weights (reference models/DeepFit.py:301-316; PointNetFeatures
// :180-208, PointNetEncoder :229-237, STN :353-380, QSTN :410-441; batch_quat_to_rotmat
// utils/normal_estimation_utils.py:365-390) with eval-mode BatchNorm folded by the host.
//
// Operand format ("tiled"): every bf16 GEMM operand -- weights and the activations this pipeline
// produces for itself -- lives in HBM in exactly the shared-memory image tcgen05.mma consumes:
// 128-row x 64-column blocks of 16 KB, inside a block 8x8 "core matrices" of 128 contiguous bytes
// ordered [k-chunk j][row-group i] (K-major, no swizzle: LBO = 2048 B between k-chunks, SBO = 128 B
// between 8-row groups).  A block is therefore ONE contiguous 16 KB TMA bulk copy
// (cp.async.bulk.shared.global + mbarrier complete_tx), needs no tensor map, and epilogue stores of
// a warp (32 consecutive rows x 16 B) are 512 contiguous bytes.
//
// Precision: the 0.01 degree parity gate needs fp32-grade products (measured: plain bf16 is ~0.2 deg
// off on the seeded network), so every operand is kept as a hi/lo bf16 pair (x = hi + lo, 16 mantissa
// bits) and each k-step issues three MMAs (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM accumulator
// (DFB_PRECISION_BF16X3).  DFB_PRECISION_BF16 issues only hi*hi.
// DFB_PRECISION_F16F8 reaches the same grade with TWO tensor-core product units per k-step instead of three: the main
// product runs on FP16 operands (11 mantissa bits, x = xh + xl), and the two first-order corrections xh*wl + xl*wh --
// 2^-12 of the main term, so four significant bits are enough for them -- run as ONE kind::f8f6f4 MMA (E4M3 operands,
// K = 32 = the k-step's 16 channels twice: activations [xh | xl*2^12], weights [wl*2^15 | wh*2^3]) at twice the FP16
// rate.  Every operand is pre-scaled by a power of two so that both MMAs accumulate 2^15 x the true product into the
// same fp32 TMEM accumulator (main plane: activations x 2^7, weights x 2^8); epilogues multiply by 2^-15.
//
// Two orientations share one mainloop:
//   N ("points are rows"): D[128 points, n] = X[128 x K] * W[n x K]^T, epilogue thread == point:
//      bias (+per-patch bias), ReLU, re-split to hi/lo bf16, 16-byte stores in tiled format.
//   T ("channels are rows"): D[128 channels, k points] = W * X^T, epilogue thread == channel:
//      the max-pool over the patch's points is a register max over TMEM columns, so the
//      [k,1024] activations of the three 128->1024 layers never leave the SM.
#include "common.cuh"
#include "../../include/deepfit_b200_debug.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <string.h>

#include <algorithm>
#include <vector>

namespace dfb {
namespace {

typedef __nv_bfloat16 bf16;

constexpr int kBlkRows = 128;
constexpr int kBlkCols = 64;
constexpr int kBlkElems = kBlkRows * kBlkCols;  // 8192 bf16 = 16 KB
constexpr int kBlkBytes = kBlkElems * 2;
constexpr uint32_t kLBO = 2048;  // bytes between adjacent 16-byte k-chunks (core matrices along K)
constexpr uint32_t kSBO = 128;   // bytes between adjacent 8-row groups (core matrices along M/N)

__host__ __device__ __forceinline__ size_t tiled_offset(long long row, int col, int KB, int br = 128) {
    // element offset of (row, col) inside a tiled matrix with KB 64-column blocks per row tile of `br`
    // rows (128, or 256 when the consumer issues N=256 MMAs: k-chunk stride LBO = br*16 bytes)
    const long long rt = row / br;
    const int r = (int)(row - rt * br);
    const int kb = col >> 6, c = col & 63;
    return ((size_t)rt * KB + kb) * ((size_t)br * 64) + (size_t)(c >> 3) * ((size_t)br * 8) + (size_t)(r >> 3) * 64 +
           (size_t)(r & 7) * 8 + (c & 7);
}

// -------------------------------------------------------------------------------------------- PTX

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a pipeline bug must never hang the GPU.  Returns false (and records `code`) on time-out.
// PARK: pass a suspend-time hint, so the hardware parks the thread until the phase completes instead of letting it
// spin through issue slots and power -- for waits that are long and not latency-critical (producers, idle warps); the
// hand-offs on the MMA <-> epilogue critical path probe without it (the parked wake-up is measurably slower).
template <bool PARK>
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    if (PARK) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(100000u)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
    return ok != 0;
}
// Bounded wait: a pipeline bug sets the error flag (dfb_model_status reports and clears it) instead of hanging the GPU.
// The bound is wall-clock time (%globaltimer, 2 s), not cycles, so GPU sharing, preemption or a profiler replay that
// stretch a wait do not trip it; the timer is only read once the first probe has failed, so a wait on a completed
// phase costs one instruction.
__device__ __forceinline__ long long wall_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <bool PARK = false>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    if (mbar_try<PARK>(bar, parity)) return true;
    const long long t0 = wall_ns();
    for (;;) {
        if (mbar_try<PARK>(bar, parity)) return true;
        const long long waited = wall_ns() - t0;
        if (waited > 2000000000LL) {
            atomicCAS(err, 0, code);
            return false;
        }
        // someone else already timed out: give up at once (persistent CTAs would otherwise stack up time-outs)
        if (waited > 1000000LL && *reinterpret_cast<volatile int*>(err) != 0) return false;
    }
}
// One lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler knows the guarded code is single-threaded
// and feeds tcgen05.mma its uniform-register operands without an ELECT / BRA.U.ANY loop around every instruction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
// A operand from tensor memory (lane = row, each 32-bit column holds two consecutive K elements), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;
}
// instruction descriptor: F32 accumulate, both operands K-major, M = 128.  Operand format code 1 = BF16 (kind::f16);
// code 0 = F16 for kind::f16 and E4M3 for kind::f8f6f4, so the FP16 main product and its FP8 correction share one
// descriptor value.
__host__ __device__ __forceinline__ uint32_t umma_idesc(int n, int fmt_code = 1) {
    return (1u << 4) | ((uint32_t)fmt_code << 7) | ((uint32_t)fmt_code << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(128 >> 4) << 24);
}
// kind::f8f6f4 (E4M3 x E4M3 -> F32, K = 32 per instruction), operands from shared memory / A from tensor memory
// (lane = row, each 32-bit column holds four consecutive K elements)
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_f8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
        : "memory");
}

// Operand formats of a model (dfb_model_desc::precision)
// FMT_F16X3: the split of FMT_BF16X3 on FP16 planes -- x * 2^7 = xh + xl, w * 2^8 = wh + wl with all four FP16 (the
// residuals reach into FP16's subnormals, spacing 2^-24 of the scaled value), three kind::f16 MMAs hi*hi + hi*lo + lo*hi:
// 22 mantissa bits per operand, i.e. fp32-grade products (the dropped lo*lo term is 2^-22 of the product) at the tensor
// cost of bf16x3, whose 16-bit operands leave the weights 1e-4 off -- enough to break the 1e-3 curvature gate on
// ill-conditioned lidar patches (oracle/studies/gate_rows_sensitivity.py).  Same plane layout as bf16x3.
enum { FMT_BF16X3 = 0, FMT_BF16 = 1, FMT_F16F8 = 2, FMT_F16X3 = 3 };
// A per-layer product count on FMT_F16X3 operands (precision ladder, DFB_PRECISION_F16MIX): same planes, same scales,
// only the hi*hi product of a k-step is issued (11-bit operands).  Never a model's format -- only ever passed to
// mma_kstep / mma_kstep_ts by a kernel whose operands are FMT_F16X3.
constexpr int FMT_F16X1 = 4;
__host__ __device__ __forceinline__ int fmt_planes(int fmt) { return fmt == FMT_BF16 ? 1 : 2; }
__host__ __device__ __forceinline__ int fmt_code(int fmt) { return (fmt == FMT_F16F8 || fmt == FMT_F16X3) ? 0 : 1; }
__host__ __device__ __forceinline__ bool fmt_is_f16(int fmt) { return fmt == FMT_F16F8 || fmt == FMT_F16X3; }
__host__ __device__ __forceinline__ bool fmt_three(int fmt) { return fmt == FMT_BF16X3 || fmt == FMT_F16X3; }
// FMT_F16F8 / FMT_F16X3 scales: main plane of activations x 2^7, of weights x 2^8; accumulators hold 2^15 x the product
constexpr float kActMain = 128.f, kWgtMain = 256.f;
__host__ __device__ __forceinline__ float fmt_acc_scale(int fmt) { return fmt_is_f16(fmt) ? (1.f / 32768.f) : 1.f; }
constexpr float kActLimit = 500.f;   // |activation| x 2^7 must stay inside FP16 (65504 / 128 = 511)

// The products of one k-step (16 channels) into accumulator d: hi*hi [+ hi*lo + lo*hi] (bf16 formats) or
// FP16 hi*hi + ONE E4M3 MMA over the correction plane (f16f8).  SS: both operands are shared-memory descriptors.
__device__ __forceinline__ void mma_kstep(int fmt, uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                          uint32_t idesc, uint32_t accum) {
    umma_bf16(d, a_hi, b_hi, idesc, accum);
    if (fmt_three(fmt)) {
        umma_bf16(d, a_hi, b_lo, idesc, 1u);
        umma_bf16(d, a_lo, b_hi, idesc, 1u);
    } else if (fmt == FMT_F16F8) {
        umma_f8(d, a_lo, b_lo, idesc, 1u);
    }
}
// TS: the A operand lives in tensor memory
__device__ __forceinline__ void mma_kstep_ts(int fmt, uint32_t d, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                             uint32_t idesc, uint32_t accum) {
    umma_bf16_ts(d, a_hi, b_hi, idesc, accum);
    if (fmt_three(fmt)) {
        umma_bf16_ts(d, a_hi, b_lo, idesc, 1u);
        umma_bf16_ts(d, a_lo, b_hi, idesc, 1u);
    } else if (fmt == FMT_F16F8) {
        umma_f8_ts(d, a_lo, b_lo, idesc, 1u);
    }
}

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// Two values at once: one packed convert for the hi parts, one for the lo parts (x = hi + lo).
// Element a lands in the low 16 bits, b in the high 16 bits (memory order a, b).
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// FMT_F16X3: scaled FP16 hi / lo pair of two values (x * scale = hi + lo, scale 2^7 activation side, 2^8 weight side)
template <bool WSIDE>
__device__ __forceinline__ void split_pair_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
    constexpr float ms = WSIDE ? 256.f : 128.f;
    const float as = a * ms, bs = b * ms;
    const __half2 h = __floats2half2_rn(as, bs);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(as - hf.x, bs - hf.y);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// the two-plane 16-bit formats (bf16x3, bf16, f16x3) behind one call
template <bool WSIDE>
__device__ __forceinline__ void split_pair_fmt(int fmt, float a, float b, uint32_t& hi, uint32_t& lo) {
    if (fmt == FMT_F16X3) split_pair_f16<WSIDE>(a, b, hi, lo);
    else split_pair(a, b, hi, lo);
}

// two floats -> two E4M3 bytes (a in the low byte, b in the high byte), round to nearest, saturating
__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(b), "f"(a));
    return (uint32_t)r;
}
// FMT_F16F8, one value pair (consecutive channels a, b) -> the packed FP16 main word and the two correction byte pairs
// c0 / c1 (16 bits each).  Activation side: main = fp16(x * 2^7), c0 = e4m3(xh), c1 = e4m3(xl * 2^12);
// weight side: main = fp16(w * 2^8), c0 = e4m3(wl * 2^15), c1 = e4m3(wh * 2^3)  (x = xh + xl, xh the FP16 value).
template <bool WSIDE>
__device__ __forceinline__ void f16f8_pair(float a, float b, uint32_t& main, uint32_t& c0, uint32_t& c1) {
    constexpr float ms = WSIDE ? kWgtMain : kActMain;
    const __half2 h = __floats2half2_rn(a * ms, b * ms);
    main = *reinterpret_cast<const uint32_t*>(&h);
    const float2 hf = __half22float2(h);
    const float ha = hf.x * (1.f / ms), hb = hf.y * (1.f / ms);
    const float la = a - ha, lb = b - hb;
    if (WSIDE) {
        c0 = e4m3x2(la * 32768.f, lb * 32768.f);
        c1 = e4m3x2(ha * 8.f, hb * 8.f);
    } else {
        c0 = e4m3x2(ha, hb);
        c1 = e4m3x2(la * 4096.f, lb * 4096.f);
    }
}
// 32 consecutive channels (starting at a multiple of 16) -> 16 main words + 16 second-plane words.  The second plane of
// a 16-channel k-step is [c0 x 16 | c1 x 16] (32 bytes), i.e. words [c0 0-15 | c1 0-15 | c0 16-31 | c1 16-31]; for the
// bf16 formats it is the lo plane.  Stored as 16-byte chunks (4 words) both planes use the same chunk index, and in
// tensor memory the same [hi 16 | second 16] column packing.  Returns the largest value seen (activation range check).
template <bool WSIDE>
__device__ __forceinline__ float split32(int fmt, const float (&f)[32], uint32_t (&ph)[16], uint32_t (&pl)[16]) {
    float mx = 0.f;
    if (fmt == FMT_F16X3) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            mx = fmaxf(mx, fmaxf(fabsf(f[2 * e]), fabsf(f[2 * e + 1])));
            split_pair_f16<WSIDE>(f[2 * e], f[2 * e + 1], ph[e], pl[e]);
        }
        return mx;
    }
    if (fmt != FMT_F16F8) {
#pragma unroll
        for (int e = 0; e < 16; ++e) split_pair(f[2 * e], f[2 * e + 1], ph[e], pl[e]);
        return mx;
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint32_t c0[8], c1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float a = f[16 * g + 2 * e], b = f[16 * g + 2 * e + 1];
            mx = fmaxf(mx, fmaxf(fabsf(a), fabsf(b)));
            f16f8_pair<WSIDE>(a, b, ph[8 * g + e], c0[e], c1[e]);
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            pl[8 * g + w] = c0[2 * w] | (c0[2 * w + 1] << 16);
            pl[8 * g + 4 + w] = c1[2 * w] | (c1[2 * w + 1] << 16);
        }
    }
    return mx;
}
// One value (row, channel) of a tiled activation matrix (the per-patch outputs of the max-pool epilogues).
__device__ __forceinline__ void store_one_act(int fmt, bf16* out_t, size_t plane, long long row, int ch, int KB, float r) {
    if (fmt == FMT_F16X3) {
        const float rs = r * kActMain;
        const __half h = __float2half_rn(rs);
        const size_t o = tiled_offset(row, ch, KB);
        reinterpret_cast<__half*>(out_t)[o] = h;
        reinterpret_cast<__half*>(out_t)[plane + o] = __float2half_rn(rs - __half2float(h));
        return;
    }
    if (fmt != FMT_F16F8) {
        bf16 hi, lo;
        split_bf16(r, hi, lo);
        const size_t o = tiled_offset(row, ch, KB);
        out_t[o] = hi;
        if (fmt == FMT_BF16X3) out_t[plane + o] = lo;
        return;
    }
    const __half h = __float2half_rn(r * kActMain);
    const float hf = __half2float(h) * (1.f / kActMain);
    reinterpret_cast<__half*>(out_t)[tiled_offset(row, ch, KB)] = h;
    // second plane: 16-channel group m = ch / 16 occupies chunks 2m (c0) and 2m + 1 (c1), one byte per channel
    const int m = ch >> 4, i = ch & 15;
    unsigned char* base = reinterpret_cast<unsigned char*>(out_t + plane);
    const uint32_t c0 = e4m3x2(hf, 0.f) & 0xffu, c1 = e4m3x2((r - hf) * 4096.f, 0.f) & 0xffu;
    base[2 * tiled_offset(row, 16 * m, KB) + i] = (unsigned char)c0;
    base[2 * tiled_offset(row, 16 * m + 8, KB) + i] = (unsigned char)c1;
}

// ------------------------------------------------------------------------------------------ GEMM

enum { EPI_ACT = 0, EPI_MAX = 1, EPI_T2 = 2, EPI_DOT = 3 };

struct GemmArgs {
    const bf16* A;       // tiled operand whose row tile is the MMA "M" side
    const bf16* B;       // tiled operand providing the N side (NT row tiles per CTA)
    size_t a_plane;      // element offset of the lo plane (hi at 0)
    size_t b_plane;
    int KB;              // 64-column k-blocks
    int fmt;             // FMT_BF16X3 / FMT_F16X3 (hi*hi + hi*lo + lo*hi), FMT_BF16 (hi*hi) or FMT_F16F8 (fp16 hi*hi + one e4m3 MMA)
    int n_mma;           // N of each UMMA (64 or 128)
    int tmode;           // 0: N orientation (A = activations), 1: T orientation (A = weights)
    int rows_per_patch;  // points per patch (N mode, activations) or 0 when rows are patches
    long long b_patch_tiles;  // N mode: B row-tile offset per patch (per-patch B operand), else 0
    const float* bias;   // N: per output column; T: per output row (channel)
    long long bias_patch_stride;  // N mode: floats between per-patch bias vectors (0 = shared)
    int relu;
    long long m_valid;   // rows (N mode) / patches (T mode) that are real
    int n_valid;         // valid output columns (N mode)
    // outputs
    bf16* out_t;         // tiled (hi plane), may be NULL
    size_t out_plane;
    int out_KB;
    int out_br;          // block rows of the tiled output (128 or 256)
    float* out_f;        // fp32 row-major, may be NULL
    long long ldo;
    // EPI_DOT
    const float* w4;
    float b4;
    int weight_mode;
    // EPI_T2
    float* trans2;       // fp32 [patch,64,64], may be NULL
    uint32_t lbo, sbo;   // descriptor strides (debug-switchable)
    int cf;              // three-product formats: keep the corrections' truncations small, see gemm_kernel (0 off, 1 on)
    int* err;
};

constexpr int kGemmThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kHeadEpiWarps = 16;   // two groups of 8 (even / odd h1 chunks)
constexpr int kHeadGroupWarps = 8;
constexpr int kHeadThreads = 96 + kHeadEpiWarps * 32;  // fused head: TMA, layer-2/3 MMA, layer-1 MMA, 16 epilogue warps

template <int NT, int EPI>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const GemmArgs g, const int stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[8], s_empty[8], s_done;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[NT * 128];
    __shared__ float s_w4[EPI == EPI_DOT ? 128 : 1];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = fmt_planes(g.fmt);
    const uint32_t stage_bytes = (uint32_t)planes * (1 + NT) * kBlkBytes;
    // tcgen05.mma accumulates in fp32 with TRUNCATION (measured: tools/tc_accum_probe.py, -0.5 ulp of the accumulator per
    // MMA): every MMA into a grown accumulator costs the same half ulp whether it adds a main product or a 2^-11 times
    // smaller correction, so a K-long three-product sum pays 3K/16 half ulps.  Kept apart from the main products the
    // corrections' truncations are 2^-11 times smaller and the sum pays K/16: the per-patch FC layers (K = 1024 / 512 /
    // 256) carried a third of the network's weight noise (oracle/studies/noise_sources.py; |dw| 2.7e-5 -> 1.9e-5 and the
    // QSTN rotation 1.0e-6 -> 4.5e-7 on the GPU, profiles/r2_fc_corrections_ab.log).  Two forms:
    //   split (NT <= 2, not the max-pool epilogue): hi*lo + lo*hi accumulate in a SECOND set of TMEM columns, added to the
    //         main accumulator (fp32, round to nearest) by the epilogue -- one pass over K, no extra operand traffic;
    //   two passes over K otherwise: the corrections of every k-step while the accumulator is still 2^-11 of its final
    //         size, then the hi*hi products.
    const bool cf = g.cf && fmt_three(g.fmt);
    const bool split = cf && NT <= 2 && EPI != EPI_MAX;
    const int passes = (cf && !split) ? 2 : 1;
    const uint32_t kTmemCols = (NT == 1 ? 128u : (NT == 2 ? 256u : 512u)) * (split ? 2u : 1u);
    const uint32_t corr_cols = (uint32_t)NT * 128u;   // split: column offset of the correction accumulators

    // tile coordinates
    long long a_tile, b_tile0, patch;
    if (g.tmode) {
        a_tile = blockIdx.x;                       // channel tile
        patch = blockIdx.y;
        b_tile0 = patch * NT;                      // the patch's NT activation row tiles
    } else {
        a_tile = blockIdx.x;                       // activation row tile
        patch = g.rows_per_patch ? (a_tile * 128) / g.rows_per_patch : 0;
        b_tile0 = (long long)blockIdx.y * NT + patch * g.b_patch_tiles;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&s_full[s]), 1);
            mbar_init(smem_u32(&s_empty[s]), 1);
        }
        mbar_init(smem_u32(&s_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&s_tmem), kTmemCols);
        tmem_relinquish();
    }
    if (warp >= 2) {
        const int t = threadIdx.x - 64;  // 0..127
        if (!g.tmode) {
            for (int c = t; c < NT * 128; c += 128) {
                const long long col = (long long)blockIdx.y * NT * 128 + c;
                s_bias[c] = (g.bias && col < g.n_valid) ? g.bias[patch * g.bias_patch_stride + col] : 0.f;
            }
        }
        if (EPI == EPI_DOT) s_w4[t] = g.w4[t];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t smem_base = smem_u32(smem);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            for (int i = 0; i < g.KB * passes; ++i) {
                const int kb = i % g.KB;
                const int s = i % stages;
                const uint32_t it = (uint32_t)(i / stages);
                if (!mbar_wait(smem_u32(&s_empty[s]), (it & 1u) ^ 1u, g.err, 1)) break;
                const uint32_t full = smem_u32(&s_full[s]);
                mbar_expect_tx(full, stage_bytes);
                uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
                for (int pl = 0; pl < planes; ++pl) {
                    bulk_g2s(dst, g.A + (size_t)pl * g.a_plane + ((size_t)a_tile * g.KB + kb) * kBlkElems, kBlkBytes, full);
                    dst += kBlkBytes;
                    for (int t = 0; t < NT; ++t) {
                        bulk_g2s(dst, g.B + (size_t)pl * g.b_plane + ((size_t)(b_tile0 + t) * g.KB + kb) * kBlkElems,
                                 kBlkBytes, full);
                        dst += kBlkBytes;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(g.n_mma, fmt_code(g.fmt));
            bool ok = true;
            for (int i = 0; i < g.KB * passes && ok; ++i) {
                const int s = i % stages;
                const uint32_t it = (uint32_t)(i / stages);
                ok = mbar_wait(smem_u32(&s_full[s]), it & 1u, g.err, 2);
                if (!ok) break;
                tc_fence_after();
                const uint32_t sa_hi = smem_base + (uint32_t)s * stage_bytes;
                const uint32_t sa_lo = sa_hi + (uint32_t)(1 + NT) * kBlkBytes;  // only valid when planes == 2
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const uint32_t sb_hi = sa_hi + (uint32_t)(1 + t) * kBlkBytes;
                    const uint32_t sb_lo = sa_lo + (uint32_t)(1 + t) * kBlkBytes;
                    const uint32_t d = tmem_base + (uint32_t)t * 128;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t koff = (uint32_t)ks * 2 * g.lbo;  // 16 elements = two 16-byte k-chunks
                        const uint32_t first = (i == 0 && ks == 0) ? 0u : 1u;
                        const uint64_t da_hi = umma_desc(sa_hi + koff, g.lbo, g.sbo), da_lo = umma_desc(sa_lo + koff, g.lbo, g.sbo);
                        const uint64_t db_hi = umma_desc(sb_hi + koff, g.lbo, g.sbo), db_lo = umma_desc(sb_lo + koff, g.lbo, g.sbo);
                        if (split) {
                            umma_bf16(d, da_hi, db_hi, idesc, first);
                            umma_bf16(d + corr_cols, da_hi, db_lo, idesc, first);
                            umma_bf16(d + corr_cols, da_lo, db_hi, idesc, 1u);
                        } else if (passes == 1) {
                            mma_kstep(g.fmt, d, da_hi, da_lo, db_hi, db_lo, idesc, first);
                        } else if (i < g.KB) {      // pass 0: hi*lo + lo*hi of every k-step
                            umma_bf16(d, da_hi, db_lo, idesc, first);
                            umma_bf16(d, da_lo, db_hi, idesc, 1u);
                        } else {                    // pass 1: hi*hi of every k-step
                            umma_bf16(d, da_hi, db_hi, idesc, 1u);
                        }
                    }
                }
                umma_commit(smem_u32(&s_empty[s]));  // frees the stage when these MMAs have read it
            }
            umma_commit(smem_u32(&s_done));  // accumulators complete
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane quadrant each) =====================
        const int q = warp & 3;
        const int lrow = q * 32 + lane;  // row of the 128-row accumulator tile owned by this thread
        const bool ok = mbar_wait(smem_u32(&s_done), 0u, g.err, 3);
        tc_fence_after();
        if (ok) {
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
            const float asc = fmt_acc_scale(g.fmt);
            uint32_t v[32];
            if (EPI == EPI_MAX) {
                float m = -3.402823466e38f;
#pragma unroll 1
                for (int c0 = 0; c0 < NT * 128; c0 += 32) {
                    tmem_ld32(tbase + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                }
                const long long ch = a_tile * 128 + lrow;
                float r = fmaf(m, asc, g.bias ? g.bias[ch] : 0.f);
                if (g.relu) r = fmaxf(r, 0.f);
                if (patch < g.m_valid) {
                    if (g.out_f) g.out_f[patch * g.ldo + ch] = r;
                    if (g.out_t) store_one_act(g.fmt, g.out_t, g.out_plane, patch, (int)ch, g.out_KB, r);
                }
            } else {
                const long long row = a_tile * 128 + lrow;
                float dot = 0.f;
#pragma unroll 1
                for (int t = 0; t < NT; ++t) {
#pragma unroll 1
                    for (int c0 = 0; c0 < g.n_mma; c0 += 32) {
                        tmem_ld32(tbase + (uint32_t)(t * 128 + c0), v);
                        if (split) {
                            uint32_t vc[32];
                            tmem_ld32(tbase + corr_cols + (uint32_t)(t * 128 + c0), vc);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vc[i]));
                        }
                        const int cl = t * 128 + c0;                                   // column inside the CTA
                        const long long col0 = (long long)blockIdx.y * NT * 128 + cl;  // global output column
                        float f[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float x = fmaf(__uint_as_float(v[i]), asc, s_bias[cl + i]);
                            if (g.relu) x = fmaxf(x, 0.f);
                            f[i] = x;
                        }
                        if (EPI == EPI_DOT) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) dot = fmaf(f[i], s_w4[cl + i], dot);
                            continue;
                        }
                        if (col0 >= g.n_valid || row >= g.m_valid) continue;
                        if (EPI == EPI_ACT) {
                            if (g.out_t) {
                                uint32_t ph[16], plo[16];
                                if (split32<false>(g.fmt, f, ph, plo) > kActLimit) atomicCAS(g.err, 0, 100);
#pragma unroll
                                for (int c8 = 0; c8 < 4; ++c8) {
                                    const size_t o = tiled_offset(row, (int)col0 + c8 * 8, g.out_KB, g.out_br);
                                    *reinterpret_cast<uint4*>(g.out_t + o) = make_uint4(ph[4 * c8], ph[4 * c8 + 1], ph[4 * c8 + 2], ph[4 * c8 + 3]);
                                    if (planes > 1)
                                        *reinterpret_cast<uint4*>(g.out_t + g.out_plane + o) =
                                            make_uint4(plo[4 * c8], plo[4 * c8 + 1], plo[4 * c8 + 2], plo[4 * c8 + 3]);
                                }
                            }
                            if (g.out_f) {
                                float4* dstf = reinterpret_cast<float4*>(g.out_f + row * g.ldo + col0);
#pragma unroll
                                for (int e = 0; e < 8; ++e) dstf[e] = make_float4(f[4 * e], f[4 * e + 1], f[4 * e + 2], f[4 * e + 3]);
                            }
                        } else if (EPI == EPI_T2) {
                            // per-patch B operand  T2^T[n][kk]  in tiled form, one 128-row block per patch.  The host
                            // permuted the fc3 rows so that output column c = (kk>>3)*512 + (n>>3)*64 + (n&7)*8 + (kk&7)
                            // is the block's own memory order: a thread's 32 columns are 64 contiguous bytes per plane
                            // T2^T is the WEIGHT-side operand of the x*T2 layer.  A thread's 32 columns are four runs of 8
                            // consecutive kk (one 16-byte chunk of the main plane each) for four rows n.  The f16f8 correction
                            // plane groups 16 kk: chunk 2m = c0 of kk 16m..16m+15, chunk 2m+1 = c1 -- a run fills one half
                            // (8 bytes) of each.
#pragma unroll
                            for (int c8 = 0; c8 < 4; ++c8) {
                                const int col = (int)col0 + c8 * 8;
                                const int n = (((col >> 6) & 7) << 3) | ((col >> 3) & 7), kk = (col >> 9) << 3;
                                const size_t o = (size_t)row * kBlkElems + (size_t)(col >> 9) * 1024 + (size_t)(col & 511);
                                uint32_t ph[4], plo[4];
                                if (g.fmt != FMT_F16F8) {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) split_pair_fmt<true>(g.fmt, f[c8 * 8 + 2 * e], f[c8 * 8 + 2 * e + 1], ph[e], plo[e]);
                                    *reinterpret_cast<uint4*>(g.out_t + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                                    if (planes > 1) *reinterpret_cast<uint4*>(g.out_t + g.out_plane + o) = make_uint4(plo[0], plo[1], plo[2], plo[3]);
                                } else {
                                    uint32_t c0[4], c1[4];
#pragma unroll
                                    for (int e = 0; e < 4; ++e) f16f8_pair<true>(f[c8 * 8 + 2 * e], f[c8 * 8 + 2 * e + 1], ph[e], c0[e], c1[e]);
                                    *reinterpret_cast<uint4*>(g.out_t + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                                    const int j = col >> 9, m2 = j >> 1, half = j & 1;
                                    unsigned char* pb = reinterpret_cast<unsigned char*>(g.out_t + g.out_plane + (size_t)row * kBlkElems);
                                    const size_t rowb = ((size_t)(n >> 3) * 64 + (size_t)(n & 7) * 8) * 2 + (size_t)half * 8;
                                    *reinterpret_cast<uint2*>(pb + (size_t)(2 * m2) * 2048 + rowb) = make_uint2(c0[0] | (c0[1] << 16), c0[2] | (c0[3] << 16));
                                    *reinterpret_cast<uint2*>(pb + (size_t)(2 * m2 + 1) * 2048 + rowb) = make_uint2(c1[0] | (c1[1] << 16), c1[2] | (c1[3] << 16));
                                }
                                if (g.trans2) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) g.trans2[(size_t)row * 4096 + (size_t)(kk + e) * 64 + n] = f[c8 * 8 + e];
                                }
                            }
                        }
                    }
                }
                if (EPI == EPI_DOT && row < g.m_valid) {
                    const float a = dot + g.b4;
                    float w;
                    if (g.weight_mode == DFB_WEIGHT_MODE_SIGMOID) w = 0.01f + 1.f / (1.f + expf(-a));
                    else w = (0.01f + 1.f + tanhf(a)) * 0.5f;
                    g.out_f[row] = w;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------- 128->1024 + max-pool, resident
// The dominant layer (63 % of the network's FLOPs).  One CTA per patch: the patch's activations
// (NT row tiles x KB k-blocks, hi+lo planes: 128 KB at k=256) are loaded ONCE and stay in shared
// memory; the weight matrix is streamed through a ring of k-block stages (A operand), every channel
// tile accumulates into one of two TMEM buffers so the max-pool epilogue of tile c overlaps the MMAs
// of tile c+1.  L2 traffic per patch: X once + W once (640 KB) instead of X x 8 + W (1.5 MB).
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t n) {   // one thread standing in for n arrivals
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(kGemmThreads) gemm_max_resident_kernel(const GemmArgs g, const int wstages) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_xfull, s_wfull[4], s_wempty[4], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = fmt_planes(g.fmt);
    const int KB = g.KB;
    const int CT = g.n_valid / 128;
    const uint32_t x_bytes = (uint32_t)planes * NT * KB * kBlkBytes;
    const uint32_t w_stage_bytes = (uint32_t)planes * kBlkBytes;
    constexpr uint32_t kTmemCols = NT == 1 ? 256 : 512;  // two accumulator buffers of NT*128 columns
    const long long patch = blockIdx.x;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&s_xfull), 1);
        for (int s = 0; s < wstages; ++s) {
            mbar_init(smem_u32(&s_wfull[s]), 1);
            mbar_init(smem_u32(&s_wempty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&s_tfull[b]), 1);
            mbar_init(smem_u32(&s_tempty[b]), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&s_tmem), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t w_base = smem_base + x_bytes;

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t xfull = smem_u32(&s_xfull);
            mbar_expect_tx(xfull, x_bytes);
            // X is stored in blocks of NT*128 rows (one patch), so each (plane, k-block) is one copy
            uint32_t dst = smem_base;
            for (int pl = 0; pl < planes; ++pl)
                for (int kb = 0; kb < KB; ++kb) {
                    bulk_g2s(dst, g.B + (size_t)pl * g.b_plane + ((size_t)patch * KB + kb) * ((size_t)NT * kBlkElems), NT * kBlkBytes, xfull);
                    dst += NT * kBlkBytes;
                }
            const int total = CT * KB;
            for (int i = 0; i < total; ++i) {
                const int s = i % wstages;
                const uint32_t it = (uint32_t)(i / wstages);
                if (!mbar_wait(smem_u32(&s_wempty[s]), (it & 1u) ^ 1u, g.err, 1)) break;
                const uint32_t full = smem_u32(&s_wfull[s]);
                mbar_expect_tx(full, w_stage_bytes);
                for (int pl = 0; pl < planes; ++pl)
                    bulk_g2s(w_base + (uint32_t)s * w_stage_bytes + (uint32_t)pl * kBlkBytes,
                             g.A + (size_t)pl * g.a_plane + (size_t)i * kBlkElems, kBlkBytes, full);  // block (ct, kb) == i
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(NT * 128, fmt_code(g.fmt));   // one MMA spans the whole patch (N = k)
            const uint32_t lbo_b = (uint32_t)NT * kLBO;
            bool ok = mbar_wait(smem_u32(&s_xfull), 0u, g.err, 2);
            for (int ct = 0; ct < CT && ok; ++ct) {
                const int buf = ct & 1;
                const uint32_t use = (uint32_t)(ct >> 1);
                if (use > 0) ok = mbar_wait(smem_u32(&s_tempty[buf]), (use - 1u) & 1u, g.err, 4);
                if (!ok) break;
                tc_fence_after();
                for (int kb = 0; kb < KB && ok; ++kb) {
                    const int i = ct * KB + kb;
                    const int s = i % wstages;
                    ok = mbar_wait(smem_u32(&s_wfull[s]), (uint32_t)(i / wstages) & 1u, g.err, 2);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t a_hi = w_base + (uint32_t)s * w_stage_bytes;
                    const uint32_t a_lo = a_hi + kBlkBytes;
                    {
                        const uint32_t b_hi = smem_base + (uint32_t)(0 * KB + kb) * (NT * kBlkBytes);
                        const uint32_t b_lo = smem_base + (uint32_t)(1 * KB + kb) * (NT * kBlkBytes);
                        const uint32_t d = tmem_base + (uint32_t)(buf * NT * 128);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t ka = (uint32_t)ks * 2 * kLBO, kb_off = (uint32_t)ks * 2 * lbo_b;
                            const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                            mma_kstep(g.fmt, d, umma_desc(a_hi + ka, kLBO, kSBO), umma_desc(a_lo + ka, kLBO, kSBO),
                                      umma_desc(b_hi + kb_off, lbo_b, kSBO), umma_desc(b_lo + kb_off, lbo_b, kSBO), idesc, first);
                        }
                    }
                    umma_commit(smem_u32(&s_wempty[s]));
                }
                umma_commit(smem_u32(&s_tfull[buf]));
            }
        }
    } else {
        const int q = warp & 3;
        const int lrow = q * 32 + lane;
        uint32_t v[32];
        for (int ct = 0; ct < CT; ++ct) {
            const int buf = ct & 1;
            const uint32_t use = (uint32_t)(ct >> 1);
            if (!mbar_wait(smem_u32(&s_tfull[buf]), use & 1u, g.err, 3)) break;
            tc_fence_after();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NT * 128);
            float mx = -3.402823466e38f;
#pragma unroll 1
            for (int c0 = 0; c0 < NT * 128; c0 += 32) {
                tmem_ld32(tbase + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_tempty[buf]));
            const long long ch = (long long)ct * 128 + lrow;
            float r = fmaf(mx, fmt_acc_scale(g.fmt), g.bias ? g.bias[ch] : 0.f);
            if (g.relu) r = fmaxf(r, 0.f);
            if (g.out_f) g.out_f[patch * g.ldo + ch] = r;
            if (g.out_t) store_one_act(g.fmt, g.out_t, g.out_plane, patch, (int)ch, g.out_KB, r);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------- fused head: 64->512 (+global bias) -> 512->256
// One CTA per 128-point row tile.  The 512-wide activation h1 = relu(W1p pf + hg[patch]) never
// leaves the SM: it is produced 64 channels at a time (MMA into one of two 64-column TMEM buffers,
// epilogue adds the per-patch bias, applies ReLU, re-splits to hi/lo bf16 and writes the chunk into
// shared memory in UMMA operand layout), and each chunk is immediately consumed as one k-block of
// the 512->256 layer, whose accumulator (256 TMEM columns) lives for the whole tile.  W1p is streamed
// in 64-row chunks (custom 8 KB blocks, LBO 1024), W2 in (k-block, n-tile) stages of 32 KB.
// Saves writing and re-reading h1 (2 KB/point) and one kernel launch.
// Optional per-CTA timeline (profiling aid, dfb_dbg_set key 7 = CTA index to trace, -1 = off): single lanes of the
// traced CTA drop clock64() stamps into a 64-slot region per kernel (chains 0-2, head 3), read by dfb_dbg_timeline.
__device__ long long g_tl_dev[4 * 64];
// dfb_dbg_set key 9 = 1: every CTA of the head kernel also records (SM id, globaltimer at entry, at exit)
__device__ long long g_cta_trace[3 * 16384];
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DFB_TL(slot) do { if (g.tl && (int)blockIdx.x == g.tl_cta) g.tl[slot] = clock64(); } while (0)

struct HeadArgs {
    const bf16* pf;      // tiled [rows, 64], hi plane; lo at +pf_plane
    size_t pf_plane;
    const bf16* w1c;     // 8 chunks x [hi 8 KB | lo 8 KB]
    const bf16* w2;      // tiled [256, 512] hi; lo at +w2_plane
    size_t w2_plane;
    const float* hg;     // fp32 [patch, 512] per-patch bias of layer 1
    const float* b2;     // fp32 [256]
    int rows_per_patch;
    long long m_valid;
    int fmt;             // FMT_*
    bf16* out_t;         // tiled [rows, 256] (only when layer 3 is not fused)
    size_t out_plane;
    int out_KB;
    // layer 3 (256->128 + ReLU) + layer 4 (128->1) + weight map, fused behind layer 2
    int fuse_l3;
    int n_tiles;         // 128-row tiles; the grid is min(n_tiles, #SMs) persistent CTAs
    const bf16* w3;      // tiled [128, 256] hi; lo at +w3_plane
    size_t w3_plane;
    const float* b3;     // fp32 [128]
    const float* w4;     // fp32 [128]
    float b4;
    // the same three vectors as kernel-parameter constants for the fused layer-3 path (no load instruction, no
    // cold-L1 miss in every CTA)
    float b2c[256];
    float b3c[128];
    float w4c[128];
    int weight_mode;
    float* out_w;        // fp32 [rows]
    long long* tl;       // timeline region or NULL
    int tl_cta;
    long long* cta_trace;  // NULL or [3 * gridDim.x]
    int dbg;             // profiling experiments only (dfb_dbg_set key 6; results are garbage): bit 1 = no layer-2 MMAs,
                         // bit 2 = no layer-2 weight loads, bit 3 = no layer-1 MMAs
    int* err;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Tail helpers of the fused head, templated on the column so that every bias / weight is a constant-bank operand.
// 32 columns of the layer-2 accumulator -> bias, ReLU, bf16 hi/lo split -> layer 3's A operand, IN PLACE: the 32 fp32
// columns become [hi 16 | lo 16] packed columns
template <int C0>
__device__ __forceinline__ void head_convert32(const HeadArgs& g, uint32_t tmem_row) {
    uint32_t v[32], ph[16], pl[16];
    float f[32];
    tmem_ld32(tmem_row + (uint32_t)C0, v);
    const float asc = fmt_acc_scale(g.fmt);
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(__uint_as_float(v[i]), asc, g.b2c[C0 + i]), 0.f);
    if (split32<false>(g.fmt, f, ph, pl) > kActLimit) atomicCAS(g.err, 0, 100);
    tmem_st16(tmem_row + (uint32_t)C0, ph);
    if (g.fmt != FMT_BF16) tmem_st16(tmem_row + (uint32_t)C0 + 16u, pl);
    tmem_st_wait();
    tc_fence_before();
}
// partial ReLU(conv3) . w4 over 32 of the 128 layer-3 channels (D3 lives at columns [256,384))
template <int C0>
__device__ __forceinline__ float head_dot32(const HeadArgs& g, uint32_t tmem_row) {
    uint32_t v[32];
    tmem_ld32(tmem_row + 256u + (uint32_t)C0, v);
    const float asc = fmt_acc_scale(g.fmt);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) dot = fmaf(fmaxf(fmaf(__uint_as_float(v[i]), asc, g.b3c[C0 + i]), 0.f), g.w4c[C0 + i], dot);
    return dot;
}

// Fused head: 64->512 (+ per-patch global bias) -> 512->256 -> 256->128 -> 128->1 -> weight map; persistent CTAs, one
// 128-row tile per iteration.  Every A operand lives in tensor memory: the tile of point features is loaded straight
// from global memory into TMEM, each layer-1 accumulator is converted IN PLACE into the bf16 hi/lo A operand of layer 2,
// and layer 2's accumulator likewise (in place) into layer 3's.  Shared memory only holds streamed weights.
// The layer-3 tail of a tile overlaps the prologue of the next one: while the layer-3 MMAs run, the (otherwise idle)
// epilogue warps store the next tile's point features and convert its first layer-1 chunk.
//   TMEM (512 columns): D2 / h2 [0,256) | D3 [256,384) | pf hi [384,416) lo [416,448) | D1/h1 buffer 0 [448,512);
//                       D1/h1 buffers 1, 2 = [256,320), [320,384) share D3's columns (chunks 1.. start after the
//                       previous tile's final dot)
//   smem (224 KB): W3A 32 KB (W3 k-block 0) | W1R 2 x 16 KB ring | W3B 64 KB (W3 k-blocks 1, 2) | W2R 3 x 32 KB ring
//                  (16 W2 stages per tile + the last W3 k-block as a 17th entry)
__global__ void __launch_bounds__(kHeadThreads) head_fused_kernel(const HeadArgs g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_pf, s_w1f[2], s_w1e[2], s_w2f[3], s_w2e[3], s_d1f[3], s_h1f[3], s_h1e[3], s_d2f, s_w3f[3], s_h2f[2], s_d3f, s_done;
    __shared__ uint32_t s_tmem;
    __shared__ float s_hg[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = fmt_planes(g.fmt);
    const int n_tiles = g.n_tiles;   // persistent CTAs: tile = blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t W3A = smem_base;
    const uint32_t W1R = smem_base + 32 * 1024;
    const uint32_t W3B = smem_base + 64 * 1024;
    const uint32_t W2R = smem_base + 128 * 1024;
    auto w3_slot = [&](int kb) { return kb == 0 ? W3A : W3B + (uint32_t)(kb - 1) * 32768u; };
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kD3 = 256, kPfHi = 384, kPfLo = 416;
    auto buf_col = [&](int b) { return b == 0 ? 448u : (b == 1 ? 256u : 320u); };
    // chunk c of a tile uses buffer c % 3: 3, 3 and 2 uses per tile; n-th use of a buffer <-> barrier parity n & 1
    auto use_index = [&](int it, int c) { return (uint32_t)(it * ((c % 3) == 2 ? 2 : 3) + c / 3); };

    // the epilogue warps fetch a tile's point features (and the per-patch bias) one tile ahead, into registers:
    // warp (q, sub) takes rows 32q..32q+31, plane sub & 1, K half sub >> 1
    uint32_t pfr[16];
    float hg_reg = 0.f;
    auto fetch_tile = [&](long long tile) {   // epilogue warps only
        const int sub = (warp - 3) >> 2;
        const int lrow = (warp & 3) * 32 + lane;
        const int plane = sub & 1, khalf = sub >> 1;
        hg_reg = g.hg[((tile * 128) / g.rows_per_patch) * 512 + (threadIdx.x - 96)];
        if (plane < planes) {
            const bf16* src = g.pf + (size_t)plane * g.pf_plane + (size_t)tile * kBlkElems + (size_t)(lrow >> 3) * 64 + (size_t)(lrow & 7) * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(khalf * 4 + j) * 1024));
                pfr[4 * j] = t.x; pfr[4 * j + 1] = t.y; pfr[4 * j + 2] = t.z; pfr[4 * j + 3] = t.w;
            }
        }
    };
    if (warp >= 3 && (int)blockIdx.x < n_tiles) fetch_tile(blockIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&s_pf), kHeadEpiWarps);
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_w1f[i]), 1); mbar_init(smem_u32(&s_w1e[i]), 1); }
        for (int i = 0; i < 3; ++i) {
            mbar_init(smem_u32(&s_w2f[i]), 1); mbar_init(smem_u32(&s_w2e[i]), 1);
            mbar_init(smem_u32(&s_d1f[i]), 1); mbar_init(smem_u32(&s_h1f[i]), kHeadGroupWarps); mbar_init(smem_u32(&s_h1e[i]), 1);
            mbar_init(smem_u32(&s_w3f[i]), 1);
        }
        mbar_init(smem_u32(&s_d2f), 1);
        for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&s_h2f[i]), kHeadEpiWarps);
        mbar_init(smem_u32(&s_d3f), 1);
        mbar_init(smem_u32(&s_done), kHeadEpiWarps);   // "the tile's last TMEM read is done" (D3 shares columns with D1 buffers 1, 2)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&s_tmem), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
#define HTL(slot) do { if (g.tl && tile == g.tl_cta) g.tl[slot] = clock64(); } while (0)

    if (warp == 0) {
        // ===================== TMA producer: weights only =====================
        if (elect_one()) {
            bool ok = true;
            auto load_w3 = [&](int kb) {
                const uint32_t bar = smem_u32(&s_w3f[kb]);
                mbar_expect_tx(bar, (uint32_t)planes * kBlkBytes);
                for (int pl = 0; pl < planes; ++pl)
                    bulk_g2s(w3_slot(kb) + (uint32_t)pl * kBlkBytes, g.w3 + (size_t)pl * g.w3_plane + (size_t)kb * kBlkElems, kBlkBytes, bar);
            };
            auto load_w1 = [&](int c) {
                const int s = c & 1;
                if (!mbar_wait<true>(smem_u32(&s_w1e[s]), ((uint32_t)(c >> 1) & 1u) ^ 1u, g.err, 1)) { ok = false; return; }
                const uint32_t bar = smem_u32(&s_w1f[s]);
                mbar_expect_tx(bar, (uint32_t)planes * 8192u);
                bulk_g2s(W1R + (uint32_t)s * 16384u, g.w1c + (size_t)c * 8192, (uint32_t)planes * 8192u, bar);  // hi|lo contiguous
            };
            int w2i = 0;   // runs across tiles: the W2 ring never stops (17 entries per tile when layer 3 is fused)
            auto load_ring = [&](const bf16* src, size_t plane_elems) {   // one [hi 16 KB | lo 16 KB] entry
                const int s = w2i % 3;
                if (!mbar_wait<true>(smem_u32(&s_w2e[s]), ((uint32_t)(w2i / 3) & 1u) ^ 1u, g.err, 1)) { ok = false; return; }
                const uint32_t bar = smem_u32(&s_w2f[s]);
                if (g.dbg & 4) {   // experiment: no layer-2 weight traffic
                    mbar_arrive(bar);
                    ++w2i;
                    return;
                }
                mbar_expect_tx(bar, (uint32_t)planes * kBlkBytes);
                for (int pl = 0; pl < planes; ++pl)
                    bulk_g2s(W2R + (uint32_t)s * 32768u + (uint32_t)pl * kBlkBytes, src + (size_t)pl * plane_elems, kBlkBytes, bar);
                ++w2i;
            };
            // W2 is packed in 256-row blocks (all 256 output channels of one 64-wide k-block = 32 KB per
            // plane); a stage is half a block: k-chunks 4h..4h+3 = 16 KB per plane, contiguous
            auto load_w2 = [&](int c, int h) { load_ring(g.w2 + (size_t)c * (2 * kBlkElems) + (size_t)h * kBlkElems, g.w2_plane); };
            for (int it = 0, tile = blockIdx.x; tile < n_tiles && ok; ++it, tile += gridDim.x) {
                // none of these rings is shared with layer 3 any more: the first chunks' weights arrive while the
                // previous tile is still in its tail
                load_w2(0, 0);
                if (ok) load_w2(0, 1);
                if (ok) load_w1(0);
                if (ok) load_w1(1);
                if (ok) load_w1(2);
                for (int c = 1; c < 8 && ok; ++c) {   // same order as the MMA warps consume
                    load_w2(c, 0);
                    if (ok) load_w2(c, 1);
                    if (ok && c + 2 < 8) load_w1(c + 2);
                    if (ok && c == 1 && g.fuse_l3) {
                        // W3A / W3B hold the previous tile's W3 until its layer 3 has completed
                        if (it > 0) ok = mbar_wait<true>(smem_u32(&s_d3f), (uint32_t)(it - 1) & 1u, g.err, 1);
                        if (ok) { load_w3(0); load_w3(1); load_w3(2); }
                    }
                }
                if (g.fuse_l3 && ok) {
                    HTL(59);
                    load_ring(g.w3 + (size_t)3 * kBlkElems, g.w3_plane);   // the last W3 k-block rides the W2 ring
                }
            }
        }
    } else if (warp == 2) {
        // ===================== layer-1 MMA issuer (its own thread: tcgen05.commit tracks per-thread groups, so
        // layer 1 runs ahead, throttled only by the three D1/h1 buffers, and never delays the layer-2 issue loop)
        if (elect_one()) {
            const uint32_t idesc64 = umma_idesc(64, fmt_code(g.fmt));
            bool ok = true;
            for (int it = 0, tile = blockIdx.x; tile < n_tiles && ok; ++it, tile += gridDim.x) {
            for (int c = 0; c < 8 && ok; ++c) {   // D1[c % 3] = pf * W1chunk(c)^T
                const int s = c & 1, b = c % 3;
                const uint32_t u = use_index(it, c);
                // chunk 0 only needs the point features (stored during the previous tile's tail) and buffer 0; the
                // other two buffers share columns with the previous tile's D3
                if (c == 0) ok = mbar_wait(smem_u32(&s_pf), (uint32_t)it & 1u, g.err, 2);
                if (c == 1 && it > 0 && ok) ok = mbar_wait(smem_u32(&s_done), (uint32_t)(it - 1) & 1u, g.err, 2);
                if (c == 0) HTL(2);
                if (ok) ok = mbar_wait(smem_u32(&s_w1f[s]), (uint32_t)(c >> 1) & 1u, g.err, 2);
                if (ok && u > 0) ok = mbar_wait(smem_u32(&s_h1e[b]), (u - 1u) & 1u, g.err, 4);   // layer 2 has read the buffer
                if (!ok) break;
                tc_fence_after();
                HTL(4 + c * 6);
                const uint32_t d = tmem_base + buf_col(b);
                const uint32_t w_hi = W1R + (uint32_t)s * 16384u, w_lo = w_hi + 8192u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (g.dbg & 8) break;   // experiment: no layer-1 MMAs
                    const uint32_t a_hi = tmem_base + kPfHi + (uint32_t)ks * 8u, a_lo = tmem_base + kPfLo + (uint32_t)ks * 8u;
                    const uint32_t kw = (uint32_t)ks * 2 * 1024u;
                    mma_kstep_ts(g.fmt, d, a_hi, a_lo, umma_desc(w_hi + kw, 1024u, kSBO), umma_desc(w_lo + kw, 1024u, kSBO), idesc64,
                                 ks ? 1u : 0u);
                }
                umma_commit(smem_u32(&s_w1e[s]));
                umma_commit(smem_u32(&s_d1f[b]));
            }
            }
        }
    } else if (warp == 1) {
        // ===================== layer-2 / layer-3 MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc256 = umma_idesc(256, fmt_code(g.fmt));
            bool ok = true;
            int w2i = 0;
            for (int it = 0, tile = blockIdx.x; tile < n_tiles && ok; ++it, tile += gridDim.x) {
            const uint32_t ip = (uint32_t)it & 1u;
            for (int c = 0; c < 8 && ok; ++c) {   // D2 += h1chunk(c) * W2[:, 64c..64c+64]^T, A = the converted D1 buffer
                const int b = c % 3;
                ok = mbar_wait(smem_u32(&s_h1f[b]), use_index(it, c) & 1u, g.err, 5);
                if (!ok) break;
                HTL(4 + c * 6 + 4);
                const uint32_t buf = tmem_base + buf_col(b);
                for (int h = 0; h < 2 && ok; ++h) {
                    const int s = w2i % 3;
                    ok = mbar_wait(smem_u32(&s_w2f[s]), (uint32_t)(w2i / 3) & 1u, g.err, 2);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t b_hi = W2R + (uint32_t)s * 32768u, b_lo = b_hi + kBlkBytes;
                    const uint32_t d = tmem_base;   // one N=256 accumulator (overwrites the previous tile's h2: the tensor
                                                    // pipe runs this thread's MMAs in order, layer 3 of that tile first)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        if (g.dbg & 2) break;
                        // k-step 2h+ks of the chunk: the epilogue half h wrote its 32 channels as [hi 16 | lo 16] columns
                        const uint32_t a_hi = buf + (uint32_t)h * 32u + (uint32_t)ks * 8u, a_lo = a_hi + 16u;
                        const uint32_t kw = (uint32_t)ks * 2 * (2 * kLBO);       // 256-row block: LBO = 4096
                        mma_kstep_ts(g.fmt, d, a_hi, a_lo, umma_desc(b_hi + kw, 2 * kLBO, kSBO), umma_desc(b_lo + kw, 2 * kLBO, kSBO),
                                     idesc256, (c == 0 && h == 0 && ks == 0) ? 0u : 1u);
                    }
                    umma_commit(smem_u32(&s_w2e[s]));
                    ++w2i;
                }
                if (ok) umma_commit(smem_u32(&s_h1e[b]));
                HTL(4 + c * 6 + 5);
            }
            if (ok) umma_commit(smem_u32(&s_d2f));
            if (ok && g.fuse_l3) {
                const uint32_t idesc128 = umma_idesc(128, fmt_code(g.fmt));
                for (int kb = 0; kb < 4 && ok; ++kb) {
                    // k-blocks 2r, 2r+1 of h2 are converted (in place, inside D2) by all 16 epilogue warps in their round
                    // r, so layer 3 starts while the second half of the layer-2 accumulator is still being converted
                    ok = mbar_wait(smem_u32(&s_h2f[kb >> 1]), ip, g.err, 5);
                    if (kb == 0) HTL(54);
                    uint32_t b_hi;
                    int s = 0;
                    if (kb < 3) {
                        ok = ok && mbar_wait(smem_u32(&s_w3f[kb]), ip, g.err, 2);
                        b_hi = w3_slot(kb);
                    } else {
                        s = w2i % 3;
                        ok = ok && mbar_wait(smem_u32(&s_w2f[s]), (uint32_t)(w2i / 3) & 1u, g.err, 2);
                        b_hi = W2R + (uint32_t)s * 32768u;
                    }
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t b_lo = b_hi + kBlkBytes;
                    const uint32_t d = tmem_base + kD3;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        // k-step ks of k-block kb: 32-channel group 2kb + (ks >> 1), [hi 16 | lo 16] columns
                        const uint32_t a_hi = tmem_base + (uint32_t)(2 * kb + (ks >> 1)) * 32u + (uint32_t)(ks & 1) * 8u, a_lo = a_hi + 16u;
                        const uint32_t koff = (uint32_t)ks * 2 * kLBO;
                        mma_kstep_ts(g.fmt, d, a_hi, a_lo, umma_desc(b_hi + koff, kLBO, kSBO), umma_desc(b_lo + koff, kLBO, kSBO), idesc128,
                                     (kb == 0 && ks == 0) ? 0u : 1u);
                    }
                    if (kb == 3) {
                        umma_commit(smem_u32(&s_w2e[s]));
                        ++w2i;
                    }
                }
                if (ok) umma_commit(smem_u32(&s_d3f));
                HTL(55);
            }
            }
        }
    } else {
        // 16 epilogue warps in two groups: group = chunk parity (even / odd 64-channel chunks of h1), inside a
        // group TMEM lane quadrant q = warp & 3 (hardware rule) and column half = bit 2 of the warp's index
        const int ew = warp - 3;
        const int q = warp & 3;
        const int grp = ew >> 3;
        const int half = (ew >> 2) & 1;
        const int quarter = ew >> 2;           // column quarter for the 256- and 128-column passes
        const int lrow = q * 32 + lane;
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        const uint32_t tmem_row = tmem_base + tlane;
        uint32_t v[32];
        const float asc = fmt_acc_scale(g.fmt);
        bool ok = true;   // after a time-out the warp keeps walking its tiles (the named barriers need all 16 warps) but
                          // skips every wait and all the work
        const bool tl_lane = lane == 0 && (ew & 7) == 0;   // one stamping lane per epilogue group
        float hg_tile = 0.f;   // this thread's element of the bias vector of the tile whose prologue ran last
        // bias + ReLU + split, in place: this warp's 32 accumulator columns of chunk c become [hi 16 | lo 16] packed columns
        auto convert_chunk = [&](int it, int tile, int c, const float* bias32, bool bias_global) {
            const int b = c % 3;
            ok = mbar_wait(smem_u32(&s_d1f[b]), use_index(it, c) & 1u, g.err, 3);
            if (!ok) return;
            tc_fence_after();
            if (tl_lane) HTL(4 + c * 6 + 1);
            const uint32_t cols = tmem_row + buf_col(b) + (uint32_t)half * 32u;
            tmem_ld32(cols, v);
            uint32_t ph[16], pl[16];
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float bi = bias_global ? __ldg(bias32 + i) : bias32[i];
                f[i] = fmaxf(fmaf(__uint_as_float(v[i]), asc, bi), 0.f);
            }
            if (split32<false>(g.fmt, f, ph, pl) > kActLimit) atomicCAS(g.err, 0, 100);
            tmem_st16(cols, ph);
            if (g.fmt != FMT_BF16) tmem_st16(cols + 16u, pl);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_h1f[b]));
            if (tl_lane) HTL(4 + c * 6 + 3);
        };
        // A tile's prologue: its point features (fetched one tile ahead) become the TMEM A operand of layer 1, and group 0
        // converts the first layer-1 chunk (bias straight from global memory: s_hg still belongs to the previous tile).
        // Runs inside the previous tile's layer-3 tail; the columns it touches ([384,512)) are free there.
        auto pro_store = [&](int it, int tile) {
            if (tl_lane && grp == 0) HTL(0);
            if (ok) {
                const int plane = quarter & 1, khalf = quarter >> 1;
                if (plane < planes) {
                    tmem_st16(tmem_row + (plane ? kPfLo : kPfHi) + (uint32_t)khalf * 16u, pfr);
                    tmem_st_wait();
                }
                tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_pf));
            hg_tile = hg_reg;
            if (tile + (int)gridDim.x < n_tiles) fetch_tile(tile + gridDim.x);   // in flight for a whole tile
        };
        auto pro_convert = [&](int it, int tile) {
            if (grp == 0 && ok)
                convert_chunk(it, tile, 0, g.hg + ((long long)tile * 128 / g.rows_per_patch) * 512 + half * 32, true);
        };
        bool pro_done = false;
        for (int it = 0, tile = blockIdx.x; tile < n_tiles; ++it, tile += gridDim.x) {
        const uint32_t ip = (uint32_t)it & 1u;
        const long long row = (long long)tile * 128 + lrow;
        if (tl_lane && grp == 0 && g.cta_trace && tile < 16384) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g.cta_trace[3 * tile] = smid;
            g.cta_trace[3 * tile + 1] = global_ns();
        }
        if (!pro_done) {
            pro_store(it, tile);
            pro_convert(it, tile);
        }
        pro_done = false;
        s_hg[threadIdx.x - 96] = hg_tile;
        asm volatile("bar.sync 1, %0;" ::"n"(kHeadEpiWarps * 32) : "memory");   // the 16 epilogue warps only
        for (int c = grp ? 1 : 2; c < 8 && ok; c += 2) convert_chunk(it, tile, c, s_hg + c * 64 + half * 32, false);
        if (ok) ok = mbar_wait(smem_u32(&s_d2f), ip, g.err, 3);
        tc_fence_after();
        if (tl_lane && grp == 0) HTL(52);
        if (!g.fuse_l3) {
            if (ok) {   // h2 goes to HBM for a separate layer-3 launch (A/B-test path)
#pragma unroll 1
                for (int c0 = quarter * 64; c0 < quarter * 64 + 64; c0 += 32) {
                    tmem_ld32(tmem_row + (uint32_t)c0, v);
                    if (row >= g.m_valid) continue;
                    uint32_t ph[16], pl[16];
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(__uint_as_float(v[i]), asc, __ldg(g.b2 + c0 + i)), 0.f);
                    split32<false>(g.fmt, f, ph, pl);
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const size_t o = tiled_offset(row, c0 + c8 * 8, g.out_KB);
                        *reinterpret_cast<uint4*>(g.out_t + o) = make_uint4(ph[4 * c8], ph[4 * c8 + 1], ph[4 * c8 + 2], ph[4 * c8 + 3]);
                        if (g.fmt != FMT_BF16)
                            *reinterpret_cast<uint4*>(g.out_t + g.out_plane + o) = make_uint4(pl[4 * c8], pl[4 * c8 + 1], pl[4 * c8 + 2], pl[4 * c8 + 3]);
                    }
                }
                tc_fence_before();
            }
        } else {
            // the next tile's point features go to TMEM first: its first layer-1 chunk then slips into the tensor pipe
            // ahead of this tile's layer 3 ([384,448) is free: every layer-1 MMA of this tile completed before D2 did)
            const bool has_next = tile + (int)gridDim.x < n_tiles;
            // h2 stays on chip as the A operand of layer 3, converted in place.  Two rounds of 128 channels (32 per warp):
            // layer 3 starts on k-blocks 0-1 while k-blocks 2-3 are still being converted.
#pragma unroll 1
            for (int r = 0; r < 2; ++r) {
                if (ok) {
                    switch (r * 4 + quarter) {
                        case 0: head_convert32<0>(g, tmem_row); break;
                        case 1: head_convert32<32>(g, tmem_row); break;
                        case 2: head_convert32<64>(g, tmem_row); break;
                        case 3: head_convert32<96>(g, tmem_row); break;
                        case 4: head_convert32<128>(g, tmem_row); break;
                        case 5: head_convert32<160>(g, tmem_row); break;
                        case 6: head_convert32<192>(g, tmem_row); break;
                        default: head_convert32<224>(g, tmem_row); break;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&s_h2f[r]));
            }
            if (tl_lane && grp == 0) HTL(53);
            // while layer 3 runs: the next tile's point features go to TMEM ([384,448) is free: every layer-1 MMA of this
            // tile completed before D2 did), its first layer-1 chunk runs between this tile's layer-3 MMAs and becomes
            // layer 2's first A operand
            if (has_next) {
                pro_store(it + 1, tile + gridDim.x);
                pro_convert(it + 1, tile + gridDim.x);
                pro_done = true;
            }
            // ReLU(conv3) . w4: every warp takes 32 of the 128 layer-3 channels of its rows, the partial sums meet
            // in shared memory (s_hg is dead by now) and the column-quarter-0 warps finish the weight map
            if (ok) ok = mbar_wait(smem_u32(&s_d3f), ip, g.err, 3);
            tc_fence_after();
            if (tl_lane && grp == 0) HTL(56);
            float dot = 0.f;
            if (ok) {
                switch (quarter) {
                    case 0: dot = head_dot32<0>(g, tmem_row); break;
                    case 1: dot = head_dot32<32>(g, tmem_row); break;
                    case 2: dot = head_dot32<64>(g, tmem_row); break;
                    default: dot = head_dot32<96>(g, tmem_row); break;
                }
                tc_fence_before();
            }
            s_hg[quarter * 128 + lrow] = dot;
            asm volatile("bar.sync 1, %0;" ::"n"(kHeadEpiWarps * 32) : "memory");   // the 16 epilogue warps only
            if (quarter == 0 && ok && row < g.m_valid) {
                const float a = ((s_hg[lrow] + s_hg[128 + lrow]) + (s_hg[256 + lrow] + s_hg[384 + lrow])) + g.b4;
                g.out_w[row] = (g.weight_mode == DFB_WEIGHT_MODE_SIGMOID) ? 0.01f + 1.f / (1.f + expf(-a))
                                                                          : (0.01f + 1.f + tanhf(a)) * 0.5f;
            }
            if (tl_lane && grp == 0) HTL(57);
        }
        // the tile's last TMEM read is done: chunks 1.. of the next tile may use the buffers that share D3's columns
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_done));
        // every warp is done with s_hg (bias vector, partial sums) before any of them writes the next tile's
        asm volatile("bar.sync 1, %0;" ::"n"(kHeadEpiWarps * 32) : "memory");
        if (tl_lane && grp == 0 && g.cta_trace && tile < 16384) g.cta_trace[3 * tile + 2] = global_ns();
        }
    }
#undef HTL
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ----------------------------------------------------------- fused chain: points -> ... -> 128 -> 1024 + max
// One CTA per patch runs a whole per-point chain of the network without touching HBM in between:
//   phase P : CUDA cores: optional rotation p' = p R, then the K=3 layer relu(W0 p' + b0) (64 channels),
//             written straight into shared memory as a tcgen05 A operand (hi/lo planes);
//   phase S : up to three 64-wide tensor-core layers (64->64, per-patch x*T2, 64->128), each
//             MMA -> TMEM -> epilogue (bias, ReLU, re-split) -> shared-memory operand of the next layer;
//             the last one writes the 128-channel activations in the 256-row block layout (N = k MMAs);
//   phase M : the resident 128->1024 + max-pool loop of gemm_max_resident_kernel over those
//             activations (weights streamed through a 3-stage ring, two TMEM accumulator buffers).
// QSTN  : P -> [64->128] -> M.   STN : P(R) -> [64->64, 64->64, 64->128] -> M.
// ENC   : P(R) -> [64->64, x*T2 (stored for the head), 64->128] -> M (no ReLU before the max).
// Shared memory (split-bf16, k=256): X 128 KB | R 96 KB;  during P/S: ACT_A = R[0,64K), ACT_B = X[0,64K), the
// three small-layer weight slots R[64K,96K), X[64K,96K), X[96K,128K); during M: R is the weight ring.  Removes pointwise3 + 6 small GEMM launches per
// forward and ~5 KB/point of HBM traffic.
struct ChainLayer {
    const bf16* w;            // tiled 128-row block(s), K = 64 (hi plane)
    size_t w_plane;           // lo plane offset (elements)
    long long w_patch_stride; // elements between per-patch operands (x*T2), 0 = shared weights
    const float* bias;        // may be NULL
    int relu;
    int n_out;                // 64 or 128
    bf16* store_t;            // optional global copy of the output (tiled, 128-row blocks, KB = 1)
    size_t store_plane;
    const float* h_bias;      // host copy of bias (launch_chain copies it into ChainArgs::biasc; unused on the device)
};

struct ChainArgs {
    const float* patch;       // f32 [n,3,k]
    const float* R;           // f32 [n,3,3] or NULL
    float* pts_out;           // optional rotated points [n,3,k]
    int k;
    // the K=3 layer and every small-layer bias travel in the kernel parameters: constant-bank operands cost no
    // load instruction and no cold-L1 miss at the start of every CTA
    float w0c[64 * 3];        // f32 [64,3]
    float b0c[64];
    float biasc[3][128];      // bias of small layer l (zeros when the layer has none)
    int n_layers;
    ChainLayer L[3];
    const bf16* w3;           // tiled [1024,128]
    size_t w3_plane;
    const float* bias3;
    int relu3;
    int n_ch;                 // 1024
    int n_patch;              // work items (patches, or 256-point halves of k=512 patches); grid = min(n_patch, #SMs) persistent CTAs
    int ipp;                  // items per patch: 1, or 2 for k = 512 (partial maxima go to out_f, chain_combine_kernel merges)
    bf16* out_t;              // tiled global-feature rows (patches)
    size_t out_plane;
    int out_KB;
    float* out_f;             // optional fp32 [n, n_ch]
    int fmt;                  // FMT_*
    int fmt_small, fmt_max;   // products of the small layers / of the max-pool layer: fmt, or FMT_F16X1 (hi*hi only) on f16x3 operands
    long long* tl;            // timeline region or NULL
    int tl_cta;
    int* err;
};

constexpr int kChainEpiWarps = 16;

// relu(W0 p + b0), all 64 channels of one point, as packed bf16 hi / lo pairs (every weight a constant-bank operand)
// (32-channel halves in the tensor-memory packing [main 16 | second plane 16])
__device__ __forceinline__ void chain_first_layer(const ChainArgs& g, float x, float y, float z, uint32_t (&ph)[32], uint32_t (&pl)[32]) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int c = 32 * hf + i;
            f[i] = fmaxf(fmaf(g.w0c[c * 3 + 2], z, fmaf(g.w0c[c * 3 + 1], y, fmaf(g.w0c[c * 3], x, g.b0c[c]))), 0.f);
        }
        if (split32<false>(g.fmt, f, reinterpret_cast<uint32_t(&)[16]>(ph[16 * hf]), reinterpret_cast<uint32_t(&)[16]>(pl[16 * hf])) > kActLimit)
            atomicCAS(g.err, 0, 100);
    }
}
constexpr int kChainThreads = 64 + kChainEpiWarps * 32;

template <int NT>
__global__ void __launch_bounds__(kChainThreads) chain_max_kernel(const ChainArgs g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_act[2], s_acc[2], s_acc_solo, s_rf[3], s_re[3], s_tf[2], s_te[2], s_b0;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = fmt_planes(g.fmt);
    const int n_patch = g.n_patch;   // persistent CTAs: patch = blockIdx.x, blockIdx.x + gridDim.x, ...
    constexpr int KB3 = 2;                                   // 128 input channels of the max layer
    constexpr uint32_t kXBytes = 2u * NT * KB3 * kBlkBytes;  // the 128-channel activations, both planes
    constexpr uint32_t kTmemCols = NT == 1 ? 256 : 512;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t XR = smem_base;             // B operand of the max-pool layer (written by the last small layer)
    const uint32_t RR = smem_base + kXBytes;   // weight ring, 3 slots of [hi 16 KB | lo 16 KB]
    const int CT = g.n_ch / 128;
    const int L = g.n_layers;
    const int TPB = CT / 2;          // channel tiles per TMEM buffer and patch
    // Activations of the small layers live in TENSOR MEMORY (A operand of tcgen05.mma), 64 columns per row tile in the
    // packed form [hi 16 | lo 16] per 32 channels: the K=3 layer writes PING, every 64-wide layer's accumulator is
    // converted in place into the next A operand (PONG, PING, ...); only the last small layer (64->128) goes to shared
    // memory, as the N=256 B operand X of the max-pool loop.  PING / PONG sit in the max-pool phase's TMEM buffer 0,
    // the last layer's 128-column accumulators in buffer 1.
    auto ping = [&](int t) { return (uint32_t)(64 * t); };
    auto pong = [&](int t) { return (uint32_t)(64 * NT + 64 * t); };
    auto lastd = [&](int t) { return (uint32_t)(128 * NT + 128 * t); };
    auto a_cols = [&](int l, int t) { return (l & 1) ? pong(t) : ping(t); };
    auto d_cols = [&](int l, int t) { return l == L - 1 ? lastd(t) : ((l & 1) ? ping(t) : pong(t)); };
    // ONE weight ring serves the whole kernel: per patch its entries are the small layers' weight blocks followed by the
    // CT * KB3 stages of W3, and the ring position keeps running across patches (slot = pos % 3, parity = (pos / 3) & 1):
    // the next patch's first-layer weights arrive while the current patch is still in its last channel tiles.
#define CTL(slot) do { if (g.tl && patch == g.tl_cta) g.tl[slot] = clock64(); } while (0)

    // phase P belongs to the eight warps that sit out the max-pool phase (column half 1); they fetch their point before
    // the set-up barrier so the DRAM latency overlaps the TMEM allocation
    float px = 0.f, py = 0.f, pz = 0.f;
    // a work item is a patch, or (k = 512) one 256-point half of a patch: item -> (patch pi, first point j0)
    const int ipp = g.ipp;
    auto fetch_point = [&](long long item) {
        const int j = (int)(item % ipp) * (NT * 128) + (((warp - 2) >> 2) & 1) * 128 + (warp & 3) * 32 + lane;
        const float* base = g.patch + (item / ipp) * 3LL * g.k;
        px = __ldg(base + j); py = __ldg(base + g.k + j); pz = __ldg(base + 2 * g.k + j);
    };
    const bool p_warp = warp >= 10 && (((warp - 2) >> 2) & 1) < NT;
    if (p_warp && (int)blockIdx.x < n_patch) fetch_point(blockIdx.x);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_act[i]), 8); mbar_init(smem_u32(&s_acc[i]), 1); }
        for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&s_rf[i]), 1); mbar_init(smem_u32(&s_re[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_tf[i]), 1); mbar_init(smem_u32(&s_te[i]), 4); }
        // "TMEM buffer 0 has been read for the last time in this patch": one phase per patch (a parity wait on s_te
        // would alias for a thread that does not follow each of its phases)
        mbar_init(smem_u32(&s_b0), 4);
        mbar_init(smem_u32(&s_acc_solo), 1);   // first layer of row tile 1 in a multi-layer chain, see phase S
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&s_tmem), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            bool ok = true;
            int pos = 0;
            // a layer that only issues hi*hi products never reads its weights' lo plane: it is not fetched
            const int planes_small = g.fmt_small == FMT_F16X1 ? 1 : planes, planes_max = g.fmt_max == FMT_F16X1 ? 1 : planes;
            auto load_entry = [&](const bf16* src, size_t plane_elems, int np) {
                const int s = pos % 3;
                ok = mbar_wait<true>(smem_u32(&s_re[s]), ((uint32_t)(pos / 3) & 1u) ^ 1u, g.err, 1);
                if (!ok) return;
                const uint32_t bar = smem_u32(&s_rf[s]);
                mbar_expect_tx(bar, (uint32_t)np * kBlkBytes);
                for (int pl = 0; pl < np; ++pl)
                    bulk_g2s(RR + (uint32_t)s * (2u * kBlkBytes) + (uint32_t)pl * kBlkBytes, src + (size_t)pl * plane_elems, kBlkBytes, bar);
                ++pos;
            };
            for (int it = 0, patch = blockIdx.x; patch < n_patch && ok; ++it, patch += gridDim.x) {
                for (int l = 0; l < L && ok; ++l) load_entry(g.L[l].w + (size_t)(patch / ipp) * g.L[l].w_patch_stride, g.L[l].w_plane, planes_small);
                for (int i = 0; i < CT * KB3 && ok; ++i) {
                    load_entry(g.w3 + (size_t)i * kBlkElems, g.w3_plane, planes_max);
                    if (i == 0) CTL(43);
                    if (i == 2) CTL(47);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            bool ok = true;
            int pos = 0;
            for (int it = 0, patch = blockIdx.x; patch < n_patch && ok; ++it, patch += gridDim.x) {
                // the small layers accumulate in TMEM buffer 0, which the previous patch's max-pool epilogue may still read
                if (it > 0) ok = mbar_wait<true>(smem_u32(&s_te[0]), (uint32_t)(it * TPB - 1) & 1u, g.err, 4);
                // the two row tiles are independent through all small layers, so they are software-pipelined: while the
                // epilogue warps of tile t turn layer l's accumulator into the next operand, the MMAs of the other tile run
                for (int l = 0; l < L && ok; ++l) {
                    const int s = pos % 3;
                    ok = mbar_wait<true>(smem_u32(&s_rf[s]), (uint32_t)(pos / 3) & 1u, g.err, 2);
                    // the last small layer accumulates in buffer 1
                    if (ok && l == L - 1 && it > 0) ok = mbar_wait<true>(smem_u32(&s_te[1]), (uint32_t)(it * TPB - 1) & 1u, g.err, 4);
                    if (!ok) break;
                    const uint32_t idesc = umma_idesc(g.L[l].n_out, fmt_code(g.fmt));
                    const uint32_t w_hi = RR + (uint32_t)s * (2u * kBlkBytes), w_lo = w_hi + kBlkBytes;
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        ok = ok && mbar_wait<true>(smem_u32(&s_act[t]), (uint32_t)l & 1u, g.err, 5);
                        if (!ok) break;
                        tc_fence_after();
                        CTL(4 + (l * 2 + t) * 3);
                        const uint32_t a = tmem_base + a_cols(l, t);
                        const uint32_t d = tmem_base + d_cols(l, t);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            // k-step ks of the 64 input channels: 32-channel half ks >> 1, [hi 16 | lo 16] columns
                            const uint32_t a_hi = a + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, a_lo = a_hi + 16u;
                            const uint32_t koff = (uint32_t)ks * 2 * kLBO;
                            mma_kstep_ts(g.fmt_small, d, a_hi, a_lo, umma_desc(w_hi + koff, kLBO, kSBO), umma_desc(w_lo + koff, kLBO, kSBO), idesc,
                                         ks ? 1u : 0u);
                        }
                        umma_commit(smem_u32((l == 0 && L > 1 && t == 1) ? &s_acc_solo : &s_acc[t]));
                    }
                    if (ok) umma_commit(smem_u32(&s_re[s]));   // the weights are dead once these MMAs have completed
                    ++pos;
                }
                for (int t = 0; t < NT && ok; ++t)
                    ok = mbar_wait<true>(smem_u32(&s_act[t]), (uint32_t)L & 1u, g.err, 5);   // X written
                CTL(22);
                const uint32_t idesc = umma_idesc(NT * 128, fmt_code(g.fmt));
                const uint32_t lbo_b = (uint32_t)NT * kLBO;
                for (int ct = 0; ct < CT && ok; ++ct) {
                    const int buf = ct & 1;
                    const uint32_t use = (uint32_t)(it * TPB + (ct >> 1));   // n-th tile into this buffer, across patches
                    if (use > 0) ok = mbar_wait<true>(smem_u32(&s_te[buf]), (use - 1u) & 1u, g.err, 4);
                    if (!ok) break;
                    tc_fence_after();
                    if (ct == 0) CTL(48);
                    for (int kb = 0; kb < KB3 && ok; ++kb) {
                        const int s = pos % 3;
                        ok = mbar_wait<true>(smem_u32(&s_rf[s]), (uint32_t)(pos / 3) & 1u, g.err, 2);
                        if (!ok) break;
                        tc_fence_after();
                        if (ct == 0 && kb == 0) CTL(23);
                        const uint32_t a_hi = RR + (uint32_t)s * (2u * kBlkBytes), a_lo = a_hi + kBlkBytes;
                        const uint32_t b_hi = XR + (uint32_t)(0 * KB3 + kb) * (NT * kBlkBytes);
                        const uint32_t b_lo = XR + (uint32_t)(1 * KB3 + kb) * (NT * kBlkBytes);
                        const uint32_t d = tmem_base + (uint32_t)(buf * NT * 128);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t ka = (uint32_t)ks * 2 * kLBO, kbo = (uint32_t)ks * 2 * lbo_b;
                            const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
                            mma_kstep(g.fmt_max, d, umma_desc(a_hi + ka, kLBO, kSBO), umma_desc(a_lo + ka, kLBO, kSBO),
                                      umma_desc(b_hi + kbo, lbo_b, kSBO), umma_desc(b_lo + kbo, lbo_b, kSBO), idesc, first);
                        }
                        umma_commit(smem_u32(&s_re[s]));
                        ++pos;
                    }
                    if (ok) umma_commit(smem_u32(&s_tf[buf]));
                    CTL(24 + (ct & 7));
                }
            }
        }
    } else {
        // ===================== 16 epilogue warps =====================
        // phase S: 4 lane quadrants x 2 row tiles x 2 column halves; phase M: the column-half-0 warps only (max over all
        // columns of a lane; the MMAs bound that phase); phase P: the column-half-1 warps, meanwhile
        const int ew = warp - 2;
        const int q = warp & 3;                   // TMEM lane quadrant (hardware rule)
        const int grp = (ew >> 2) & 1;            // phases P/S: row tile; phase M: channel-tile parity
        const int half = ew >> 3;                 // phase S: column half
        const int lrow = q * 32 + lane;
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        const bool active = grp < NT;             // owns a row of the patch in phases P/S
        const int k = g.k;
        const float asc = fmt_acc_scale(g.fmt);
        bool ok = true;
        uint32_t v[32];
        const bool tl_lane = lane == 0 && q == 0 && half == 0;   // one stamping lane per epilogue group
        // element offset of (row lrow, 8-channel chunk j) inside a 128-row operand block
        const uint32_t row_off = (uint32_t)(lrow >> 3) * 128u + (uint32_t)(lrow & 7) * 16u;
        for (int it = 0, patch = blockIdx.x; patch < n_patch && ok; ++it, patch += gridDim.x) {
        if (tl_lane && grp == 0) CTL(0);
        // ---- phase P ----
        // the first layer of the NEXT patch is computed while the tensor core is still busy with this one; only the
        // tensor-memory stores wait until buffer 0 has been read for the last time
        const long long pi = patch / ipp;         // the patch this item belongs to
        const int j0 = (patch % ipp) * (NT * 128);   // its first point
        if (p_warp) {
            const int j = j0 + grp * 128 + lrow;       // point index inside the patch
            float x = px, y = py, z = pz;
            if (patch + (int)gridDim.x < n_patch) fetch_point(patch + gridDim.x);   // next patch's point, in flight all along
            if (g.R) {
                const float* r = g.R + pi * 9;
                const float xr = fmaf(z, r[6], fmaf(y, r[3], x * r[0]));
                const float yr = fmaf(z, r[7], fmaf(y, r[4], x * r[1]));
                const float zr = fmaf(z, r[8], fmaf(y, r[5], x * r[2]));
                x = xr; y = yr; z = zr;
                if (g.pts_out) {
                    float* po = g.pts_out + pi * 3LL * k;
                    po[j] = x; po[k + j] = y; po[2 * k + j] = z;
                }
            }
            uint32_t ph[32], pl[32];
            chain_first_layer(g, x, y, z, ph, pl);
            if (it > 0) ok = mbar_wait<true>(smem_u32(&s_b0), (uint32_t)(it - 1) & 1u, g.err, 3);
            if (!ok) break;
            tc_fence_after();
            const uint32_t cols = tmem_base + tlane + ping(grp);
            tmem_st16(cols, reinterpret_cast<const uint32_t(&)[16]>(ph[0]));
            tmem_st16(cols + 32u, reinterpret_cast<const uint32_t(&)[16]>(ph[16]));
            if (planes > 1) {
                tmem_st16(cols + 16u, reinterpret_cast<const uint32_t(&)[16]>(pl[0]));
                tmem_st16(cols + 48u, reinterpret_cast<const uint32_t(&)[16]>(pl[16]));
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_n(smem_u32(&s_act[grp]), 2u);   // the barrier counts the 8 warps of the S-phase epilogues
            if (lane == 0 && q == 0) CTL(2 + grp);
        }
        // ---- phase S ----
        for (int l = 0; l < L && ok && active; ++l) {
            const ChainLayer& Ly = g.L[l];
            const bool last = (l == L - 1);
            // The warps that drained the previous patch's LAST channel tile (row tile 1, column half 0) come late to a
            // new patch.  In a multi-layer chain its first layer does not wait for them: the column-half-1 warps of tile 1
            // convert all 64 columns and arrive for both halves.
            // That accumulator has a barrier of its own (one phase per patch), so that the late warps never have to
            // follow a phase they take no part in: a parity wait two phases behind would alias.
            const bool solo = (l == 0 && L > 1 && grp == 1);
            if (solo && half == 0) continue;
            const long long grow = pi * k + j0 + grp * 128 + lrow;   // global row (point) index
            const int cbeg = solo ? 0 : half * (Ly.n_out >> 1);
            // everything that does not depend on the accumulator happens before the wait
            float bias_r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) bias_r[i] = g.biasc[l][cbeg + i];
            const float floor_v = Ly.relu ? 0.f : -3.402823466e38f;
            if (solo) ok = mbar_wait<true>(smem_u32(&s_acc_solo), (uint32_t)it & 1u, g.err, 3);
            else if (grp == 1 && L > 1) ok = mbar_wait<true>(smem_u32(&s_acc[1]), (uint32_t)(it * (L - 1) + l - 1) & 1u, g.err, 3);
            else ok = mbar_wait<true>(smem_u32(&s_acc[grp]), (uint32_t)(it * L + l) & 1u, g.err, 3);
            if (!ok) break;
            tc_fence_after();
            if (tl_lane) CTL(4 + (l * 2 + grp) * 3 + 1);
            const uint32_t dcol = tmem_base + tlane + d_cols(l, grp);
            if (!last) {
                // 64-wide layer: this warp's 32 accumulator columns become [hi 16 | lo 16] packed columns, in place
#pragma unroll 1
                for (int cb = cbeg; cb < (solo ? 64 : cbeg + 32); cb += 32) {
                    uint32_t ph[16], pl[16];
                    float f[32];
                    tmem_ld32(dcol + (uint32_t)cb, v);
                    if (tl_lane && l == 0 && grp == 0) CTL(44);
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(__uint_as_float(v[i]), asc, bias_r[i]), floor_v);
                    if (split32<false>(g.fmt, f, ph, pl) > kActLimit) atomicCAS(g.err, 0, 100);
                    tmem_st16(dcol + (uint32_t)cb, ph);
                    if (planes > 1) tmem_st16(dcol + (uint32_t)cb + 16u, pl);
                    if (Ly.store_t) {   // x * T2 is also the head's input
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            const size_t go = tiled_offset(grow, cb + c8 * 8, 1);
                            *reinterpret_cast<uint4*>(Ly.store_t + go) = make_uint4(ph[4 * c8], ph[4 * c8 + 1], ph[4 * c8 + 2], ph[4 * c8 + 3]);
                            if (planes > 1)
                                *reinterpret_cast<uint4*>(Ly.store_t + Ly.store_plane + go) = make_uint4(pl[4 * c8], pl[4 * c8 + 1], pl[4 * c8 + 2], pl[4 * c8 + 3]);
                        }
                    }
                    if (solo && cb == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) bias_r[i] = g.biasc[l][32 + i];
                    }
                }
                tmem_st_wait();
                if (tl_lane && l == 0 && grp == 0) CTL(45);
            } else {
                // last small layer (64 -> 128): X in shared memory, [plane][kb][NT*128 rows], k-chunk stride NT*2048
                const uint32_t obase = XR + (uint32_t)grp * 2048u + row_off;
                const uint32_t lo_off = (uint32_t)KB3 * NT * kBlkBytes;
                const bool x_lo = planes > 1 && g.fmt_max != FMT_F16X1;   // a hi*hi-only max-pool layer never reads X's lo plane
#pragma unroll 1
                for (int c0 = cbeg; c0 < cbeg + (Ly.n_out >> 1); c0 += 32) {
                    tmem_ld32(dcol + (uint32_t)c0, v);
                    {
                        uint32_t ph[16], pl[16];
                        float f[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(__uint_as_float(v[i]), asc, bias_r[i]), floor_v);
                        if (split32<false>(g.fmt, f, ph, pl) > kActLimit) atomicCAS(g.err, 0, 100);
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            const int col = c0 + c8 * 8;
                            const uint32_t o = obase + (uint32_t)(col >> 6) * (NT * kBlkBytes) + (uint32_t)((col & 63) >> 3) * (NT * 2048u);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o), "r"(ph[4 * c8]), "r"(ph[4 * c8 + 1]), "r"(ph[4 * c8 + 2]), "r"(ph[4 * c8 + 3]) : "memory");
                            if (x_lo)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(o + lo_off), "r"(pl[4 * c8]), "r"(pl[4 * c8 + 1]), "r"(pl[4 * c8 + 2]), "r"(pl[4 * c8 + 3]) : "memory");
                        }
                    }
                    if (c0 + 32 < cbeg + (Ly.n_out >> 1)) {   // the second 32 columns' bias
#pragma unroll
                        for (int i = 0; i < 32; ++i) bias_r[i] = g.biasc[l][c0 + 32 + i];
                    }
                }
                fence_async_smem();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_n(smem_u32(&s_act[grp]), solo ? 2u : 1u);
            if (tl_lane || (solo && lane == 0 && q == 0)) CTL(4 + (l * 2 + grp) * 3 + 2);
        }
        // ---- phase M: group g takes the channel tiles of parity g ----
        for (int ct = grp; ct < CT && ok && half == 0; ct += 2) {
            const int buf = ct & 1;                 // == grp
            const uint32_t use = (uint32_t)(it * TPB + (ct >> 1));
            ok = mbar_wait<true>(smem_u32(&s_tf[buf]), use & 1u, g.err, 3);
            if (!ok) break;
            tc_fence_after();
            if (tl_lane) CTL(32 + (ct & 7));
            float mx = -3.402823466e38f;
#pragma unroll 1
            for (int c0 = 0; c0 < NT * 128; c0 += 32) {
                tmem_ld32(tmem_base + tlane + (uint32_t)(buf * NT * 128 + c0), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&s_te[buf]));
                if (buf == 0 && ct + 2 >= CT) mbar_arrive(smem_u32(&s_b0));   // buffer 0's last tile of the patch
            }
            const long long ch = (long long)ct * 128 + lrow;
            float r = fmaf(mx, asc, g.bias3 ? g.bias3[ch] : 0.f);
            if (g.relu3) r = fmaxf(r, 0.f);
            if (g.out_f) g.out_f[(long long)patch * g.n_ch + ch] = r;
            if (g.out_t) store_one_act(g.fmt, g.out_t, g.out_plane, patch, (int)ch, g.out_KB, r);
        }
        if (tl_lane) CTL(40 + grp);
        }
    }
#undef CTL
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// k = 512: the chain kernel ran on 256-point halves (bias and ReLU are monotone, so they commute with the max);
// the global feature of a patch is the max of its two partial vectors, re-split into the tiled bf16 planes.
__global__ void chain_combine_kernel(const float* __restrict__ part, long long n_patch, int n_ch, int ipp, bf16* __restrict__ out_t,
                                     size_t out_plane, int out_KB, int fmt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_patch * n_ch) return;
    const long long p = i / n_ch;
    const int ch = (int)(i % n_ch);
    float r = part[(p * ipp) * n_ch + ch];
    for (int h = 1; h < ipp; ++h) r = fmaxf(r, part[(p * ipp + h) * n_ch + ch]);
    store_one_act(fmt, out_t, out_plane, p, ch, out_KB, r);
}

// --------------------------------------------------------------------------- small CUDA-core kernels

// fp32 row-major [rows, cols] (ld) -> tiled planes in format `fmt` (weight side or activation side for f16f8);
// rows/cols beyond the source are zero.
__global__ void pack_tiled_kernel(const float* __restrict__ src, long long rows, int cols, long long ld,
                                  long long rows_pad, int cols_pad, bf16* __restrict__ dst, size_t plane, int fmt, int wside, int br) {
    const long long total = rows_pad * (cols_pad / 8);
    const int KB = cols_pad / 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i % rows_pad;  // consecutive threads -> consecutive rows -> contiguous 16 B chunks
        const int c8 = (int)(i / rows_pad);
        uint32_t ph[4], pl[4], c0[4], c1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a = 0.f, b = 0.f;
            const int c = c8 * 8 + 2 * e;
            if (r < rows && c < cols) a = src[r * ld + c];
            if (r < rows && c + 1 < cols) b = src[r * ld + c + 1];
            if (fmt == FMT_F16X3) {
                if (wside) split_pair_f16<true>(a, b, ph[e], pl[e]);
                else split_pair_f16<false>(a, b, ph[e], pl[e]);
            } else if (fmt != FMT_F16F8) split_pair(a, b, ph[e], pl[e]);
            else if (wside) f16f8_pair<true>(a, b, ph[e], c0[e], c1[e]);
            else f16f8_pair<false>(a, b, ph[e], c0[e], c1[e]);
        }
        const size_t o = tiled_offset(r, c8 * 8, KB, br);
        *reinterpret_cast<uint4*>(dst + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        if (fmt_three(fmt)) {
            *reinterpret_cast<uint4*>(dst + plane + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        } else if (fmt == FMT_F16F8) {
            // 16-channel group m = c8 / 2: chunk 2m holds c0, chunk 2m + 1 holds c1; this thread fills 8 bytes of each
            const int m = c8 >> 1, half = c8 & 1;
            unsigned char* pb = reinterpret_cast<unsigned char*>(dst + plane);
            *reinterpret_cast<uint2*>(pb + 2 * tiled_offset(r, 16 * m, KB, br) + 8 * half) = make_uint2(c0[0] | (c0[1] << 16), c0[2] | (c0[3] << 16));
            *reinterpret_cast<uint2*>(pb + 2 * tiled_offset(r, 16 * m + 8, KB, br) + 8 * half) = make_uint2(c1[0] | (c1[1] << 16), c1[2] | (c1[3] << 16));
        }
    }
}

__device__ __forceinline__ float e4m3_to_float(unsigned char b) {
    const __half_raw hr = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)b, __NV_E4M3);
    return __half2float(*reinterpret_cast<const __half*>(&hr));
}
// tiled planes -> fp32 row-major, for tests (activation-side operands)
__global__ void unpack_tiled_kernel(const bf16* __restrict__ src, size_t plane, int fmt, long long rows, int cols,
                                    int KB, float* __restrict__ dst, int br) {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i % cols);
        const size_t o = tiled_offset(r, c, KB, br);
        float v;
        if (fmt == FMT_F16X3) {
            v = (__half2float(reinterpret_cast<const __half*>(src)[o]) + __half2float(reinterpret_cast<const __half*>(src)[plane + o])) *
                (1.f / kActMain);
        } else if (fmt != FMT_F16F8) {
            v = __bfloat162float(src[o]);
            if (fmt == FMT_BF16X3) v += __bfloat162float(src[plane + o]);
        } else {
            // x = xh + xl with xh = main / 2^7 and xl ~ c1 / 2^12 (4 significant bits of the residual)
            v = __half2float(reinterpret_cast<const __half*>(src)[o]) * (1.f / kActMain);
            const unsigned char* pb = reinterpret_cast<const unsigned char*>(src + plane);
            v += e4m3_to_float(pb[2 * tiled_offset(r, 16 * (c >> 4) + 8, KB, br) + (c & 15)]) * (1.f / 4096.f);
        }
        dst[i] = v;
    }
}

// K=3 first layers on CUDA cores: optional rotation P' = P R (DeepFit.py:188-190), then
// relu(W p + b) for 64 channels, written as a tiled hi/lo operand.  One thread per point.
__global__ void __launch_bounds__(128) pointwise3_kernel(const float* __restrict__ patch, const float* __restrict__ R,
                                                         long long n_patch, int k, const float* __restrict__ W,
                                                         const float* __restrict__ bvec, bf16* __restrict__ out,
                                                         size_t plane, int fmt, float* __restrict__ pts_out) {
    __shared__ float sW[64 * 3 + 64];
    for (int i = threadIdx.x; i < 64 * 3; i += blockDim.x) sW[i] = W[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) sW[192 + i] = bvec[i];
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = n_patch * k;
    if (row >= total) return;
    const long long p = row / k;
    const int j = (int)(row - p * k);
    const float* base = patch + p * 3LL * k;
    float x = base[j], y = base[k + j], z = base[2 * k + j];
    if (R) {
        const float* r = R + p * 9;
        const float xr = fmaf(z, r[6], fmaf(y, r[3], x * r[0]));
        const float yr = fmaf(z, r[7], fmaf(y, r[4], x * r[1]));
        const float zr = fmaf(z, r[8], fmaf(y, r[5], x * r[2]));
        x = xr; y = yr; z = zr;
        if (pts_out) {
            float* po = pts_out + p * 3LL * k;
            po[j] = x; po[k + j] = y; po[2 * k + j] = z;
        }
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        uint32_t ph[16], pl[16];
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int c = 32 * hf + i;
            f[i] = fmaxf(fmaf(sW[c * 3 + 2], z, fmaf(sW[c * 3 + 1], y, fmaf(sW[c * 3], x, sW[192 + c]))), 0.f);
        }
        split32<false>(fmt, f, ph, pl);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            const size_t o = tiled_offset(row, 32 * hf + c8 * 8, 1);
            *reinterpret_cast<uint4*>(out + o) = make_uint4(ph[4 * c8], ph[4 * c8 + 1], ph[4 * c8 + 2], ph[4 * c8 + 3]);
            if (fmt != FMT_BF16) *reinterpret_cast<uint4*>(out + plane + o) = make_uint4(pl[4 * c8], pl[4 * c8 + 1], pl[4 * c8 + 2], pl[4 * c8 + 3]);
        }
    }
}

// QSTN tail: q = fc3(f) + (1,0,0,0); R = quat_to_rotmat(q)  (DeepFit.py:432-439,
// utils/normal_estimation_utils.py:365-390).  One warp per patch.
__global__ void quat_kernel(const float* __restrict__ f, long long n, const float* __restrict__ W, const float* __restrict__ b,
                            float* __restrict__ R) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n) return;
    float q[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < 256; j += 32) {
        const float x = f[p * 256 + j];
#pragma unroll
        for (int c = 0; c < 4; ++c) q[c] = fmaf(W[c * 256 + j], x, q[c]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) q[c] = warp_sum(q[c]) + b[c];
    if (lane == 0) {
        q[0] += 1.f;
        const float s = 2.f / (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        float* o = R + p * 9;
        o[0] = 1.f - (q[2] * q[2] + q[3] * q[3]) * s;
        o[1] = (q[1] * q[2] - q[3] * q[0]) * s;
        o[2] = (q[1] * q[3] + q[2] * q[0]) * s;
        o[3] = (q[1] * q[2] + q[3] * q[0]) * s;
        o[4] = 1.f - (q[1] * q[1] + q[3] * q[3]) * s;
        o[5] = (q[2] * q[3] - q[1] * q[0]) * s;
        o[6] = (q[1] * q[3] - q[2] * q[0]) * s;
        o[7] = (q[2] * q[3] + q[1] * q[0]) * s;
        o[8] = 1.f - (q[1] * q[1] + q[2] * q[2]) * s;
    }
}

// ------------------------------------------------------------------------------------ host side

int g_swap_lbo_sbo = 0;  // debug switch (dfb_dbg_set)

template <int NT, int EPI>
int launch_gemm_t(GemmArgs g, dim3 grid, cudaStream_t st) {
    const int planes = fmt_planes(g.fmt);
    const size_t stage_bytes = (size_t)planes * (1 + NT) * kBlkBytes;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > g.KB) stages = g.KB;
    if (stages > 8) stages = 8;
    if (stages < 1) stages = 1;
    const size_t smem = stages * stage_bytes;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_kernel<NT, EPI>), 200 * 1024)) return rc;
    g.lbo = g_swap_lbo_sbo ? kSBO : kLBO;
    g.sbo = g_swap_lbo_sbo ? kLBO : kSBO;
    gemm_kernel<NT, EPI><<<grid, kGemmThreads, smem, st>>>(g, stages);
    note_launch();
    return check_cuda(cudaGetLastError(), "gemm_kernel launch", __FILE__, __LINE__);
}

int launch_gemm(const GemmArgs& g, int NT, int epi, dim3 grid, cudaStream_t st) {
#define DFB_CASE(nt, ep) if (NT == nt && epi == ep) return launch_gemm_t<nt, ep>(g, grid, st)
    DFB_CASE(1, EPI_ACT); DFB_CASE(2, EPI_ACT);
    DFB_CASE(1, EPI_MAX); DFB_CASE(2, EPI_MAX); DFB_CASE(3, EPI_MAX); DFB_CASE(4, EPI_MAX);
    DFB_CASE(1, EPI_T2); DFB_CASE(1, EPI_DOT);
#undef DFB_CASE
    set_error("launch_gemm: unsupported configuration NT=%d epi=%d", NT, epi);
    return DFB_E_INTERNAL;
}

struct TiledBuf {
    bf16* p = nullptr;
    size_t plane = 0;  // elements per plane
    int KB = 0;
    long long rows_pad = 0;
    int br = 128;      // rows per block
};

int alloc_tiled(TiledBuf& t, long long rows, int cols, int br = 128) {
    t.br = br;
    t.rows_pad = (rows + br - 1) / br * br;
    t.KB = (cols + 63) / 64;
    t.plane = (size_t)t.rows_pad * t.KB * 64;
    DFB_CUDA(cudaMalloc(&t.p, t.plane * 2 * sizeof(bf16)));
    return 0;
}

struct PackedLayer {
    TiledBuf w;        // tiled weights [out_pad, in_pad]
    float* bias = nullptr;  // device fp32 [out]
    std::vector<float> h_bias;  // host copy (kernel-parameter constants of the fused chain kernels)
    int in = 0, out = 0;
};

}  // namespace
}  // namespace dfb

struct dfb_model {
    int k = 0, weight_mode = 0, fmt = 0 /* FMT_* */, max_batch = 0;
    int x1_mask = 0;  // DFB_PRECISION_F16MIX: chain layers that issue only the hi*hi product (DFB_MIX_* bits)
    float b4 = 0.f;  // conv4 bias
    dfb::PackedLayer L[DFB_NUM_LAYERS];
    dfb::PackedLayer head_global;  // HEAD_CONV1 columns 0..1023  (bias = folded conv1 bias)
    dfb::PackedLayer head_point;   // HEAD_CONV1 columns 1024..1087 (no bias of its own)
    float* w_small[DFB_NUM_LAYERS] = {nullptr};  // raw fp32 weights of the CUDA-core layers
    float* chain_part = nullptr;   // k = 512: fp32 [2 * max_batch, 1024] partial maxima of the half-patch work items
    std::vector<float> h_w_small[DFB_NUM_LAYERS];  // host copies of the two K=3 layers and of conv4
    dfb::bf16* head_w1c = nullptr;               // HEAD_CONV1 point part in 64-row chunks for the fused head
    dfb::PackedLayer head_w2_256;                // HEAD_CONV2 in 256-row blocks (N=256 MMAs of the fused head)
    // workspace
    dfb::TiledBuf x64a, x64b, x64c, x128, h512, h256, gfeat, f512, f256, t2blk;
    float* f256_f = nullptr;  // fp32 [max_batch,256]
    float* hg = nullptr;      // fp32 [max_batch,512]
    int* err = nullptr;       // device error flag
    std::vector<void*> allocs;
    // optional per-kernel-class timing
    struct ProfRec { int cls; cudaEvent_t e0, e1; };
    bool prof = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
};

namespace dfb {
namespace {
// Activation buffers of the unfused (debug / k=384,512) paths, allocated when first needed.
int ensure_unfused_buffers(dfb_model* m) {
    if (m->x64a.p) return 0;
    const long long R = (long long)m->max_batch * m->k;
    TiledBuf* bufs[] = {&m->x64a, &m->x64b, &m->x128, &m->h512, &m->h256};
    const int widths[] = {64, 64, 128, 512, 256};
    for (int i = 0; i < 5; ++i) {
        // x128 feeds the resident max-pool kernel, which issues N=256 MMAs over 256-row blocks at k=256
        int rc = alloc_tiled(*bufs[i], R, widths[i], (bufs[i] == &m->x128 && m->k == 256) ? 256 : 128);
        if (rc) return rc;
        m->allocs.push_back(bufs[i]->p);
    }
    return 0;
}

struct ProfScope {
    dfb_model* m;
    int cls;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t get(dfb_model* m) {
        if (!m->pool.empty()) {
            cudaEvent_t e = m->pool.back();
            m->pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    ProfScope(dfb_model* m_, int cls_, cudaStream_t st_) : m(m_), cls(cls_), st(st_) {
        if (m->prof) {
            e0 = get(m);
            e1 = get(m);
            cudaEventRecord(e0, st);
        }
    }
    ~ProfScope() {
        if (m->prof) {
            cudaEventRecord(e1, st);
            m->recs.push_back({cls, e0, e1});
        }
    }
};
}  // namespace
}  // namespace dfb

namespace dfb {
namespace {

// Cluster of 2 with TMA multicast of the weight stages: correct and tested, but measured SLOWER than plain
// launches (head 47.7 -> 41.7 ms per 131 072 patches without it): B200's L2 already de-duplicates
// same-line requests from <= 4 SMs, so multicast at cluster size 2 saves no L2 bandwidth and only adds the
// lock-step of the pair.  Off by default; dfb_dbg_set key 3 turns it on.

int g_tl_cta = -1;   // dfb_dbg_set key 7: CTA whose timeline is recorded (-1 = off)
long long* timeline_region(int region) {
    if (g_tl_cta < 0) return nullptr;
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_tl_dev) != cudaSuccess) return nullptr;
    return static_cast<long long*>(p) + region * 64;
}
int g_cta_trace_on = 0;   // dfb_dbg_set key 9
int g_head_persistent = 1;  // dfb_dbg_set key 8: 0 = one CTA per tile
int g_head_dbg = 0;  // profiling experiments (dfb_dbg_set key 6), see HeadArgs::dbg
int g_head_l3 = 1;  // debug switch (dfb_dbg_set key 4): 0 = layer 3 + weight epilogue as a separate launch

int launch_head_fused(const dfb_model* m, const TiledBuf& pf, long long rows, const float* hg, const PackedLayer& W2,
                      TiledBuf& out, float* out_w, cudaStream_t st) {
    HeadArgs a;
    memset(&a, 0, sizeof(a));
    a.pf = pf.p; a.pf_plane = pf.plane;
    a.w1c = m->head_w1c;
    a.w2 = W2.w.p; a.w2_plane = W2.w.plane;
    a.hg = hg; a.b2 = W2.bias;
    a.rows_per_patch = m->k;
    a.m_valid = rows;
    a.fmt = m->fmt;
    a.out_t = out.p; a.out_plane = out.plane; a.out_KB = out.KB;
    a.fuse_l3 = out_w != nullptr;
    a.w3 = m->L[DFB_L_HEAD_CONV3].w.p; a.w3_plane = m->L[DFB_L_HEAD_CONV3].w.plane;
    a.b3 = m->L[DFB_L_HEAD_CONV3].bias;
    a.w4 = m->w_small[DFB_L_HEAD_CONV4];
    memcpy(a.b2c, m->L[DFB_L_HEAD_CONV2].h_bias.data(), sizeof(a.b2c));
    memcpy(a.b3c, m->L[DFB_L_HEAD_CONV3].h_bias.data(), sizeof(a.b3c));
    memcpy(a.w4c, m->h_w_small[DFB_L_HEAD_CONV4].data(), sizeof(a.w4c));
    a.b4 = m->b4;
    a.weight_mode = m->weight_mode;
    a.out_w = out_w;
    a.dbg = g_head_dbg;
    if (g_cta_trace_on) {
        void* tp = nullptr;
        if (cudaGetSymbolAddress(&tp, g_cta_trace) == cudaSuccess) a.cta_trace = static_cast<long long*>(tp);
    }
    a.tl = timeline_region(3); a.tl_cta = g_tl_cta;
    a.err = m->err;
    const size_t smem = 224 * 1024;
    const unsigned tiles = (unsigned)((rows + 127) / 128);
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&head_fused_kernel), (int)smem)) return rc;
    a.n_tiles = (int)tiles;
    const unsigned grid = g_head_persistent ? std::min<unsigned>(tiles, (unsigned)sm_count()) : tiles;
    head_fused_kernel<<<grid, kHeadThreads, smem, st>>>(a);
    note_launch();
    return check_cuda(cudaGetLastError(), "head_fused_kernel launch", __FILE__, __LINE__);
}

int upload_pack(int fmt, const float* h_w, int out, int in, int col0, int ncols, long long ldw, PackedLayer& pl, cudaStream_t st,
                const int* row_perm = nullptr, int br = 128) {
    // host fp32 [out, ldw] (columns col0..col0+ncols) -> device tiled hi/lo
    std::vector<float> tmp((size_t)out * ncols);
    for (int r = 0; r < out; ++r) {
        const int src = row_perm ? row_perm[r] : r;
        for (int c = 0; c < ncols; ++c) tmp[(size_t)r * ncols + c] = h_w[(size_t)src * ldw + col0 + c];
    }
    float* d_tmp = nullptr;
    DFB_CUDA(cudaMalloc(&d_tmp, tmp.size() * sizeof(float)));
    DFB_CUDA(cudaMemcpyAsync(d_tmp, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = alloc_tiled(pl.w, out, ncols, br);
    if (rc) return rc;
    pl.in = ncols;
    pl.out = out;
    const long long total = pl.w.rows_pad * (pl.w.KB * 8);
    pack_tiled_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_tmp, out, ncols, ncols, pl.w.rows_pad, pl.w.KB * 64,
                                                                        pl.w.p, pl.w.plane, fmt, 1, pl.w.br);
    note_launch();
    DFB_CUDA(cudaGetLastError());
    DFB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_tmp);
    return 0;
}

int upload_f32(const float* h, size_t n, float** d, cudaStream_t st) {
    DFB_CUDA(cudaMalloc(d, n * sizeof(float)));
    DFB_CUDA(cudaMemcpyAsync(*d, h, n * sizeof(float), cudaMemcpyHostToDevice, st));
    DFB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int g_gemm_cf = 1;  // dfb_dbg_set key 11: 0 = one pass over K, products in hh, hl, lh order per k-step

GemmArgs base_args(const dfb_model* m) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.fmt = m->fmt;
    g.cf = g_gemm_cf;
    g.n_mma = 128;
    g.err = m->err;
    g.weight_mode = m->weight_mode;
    g.b4 = m->b4;
    return g;
}

// N orientation: out = act(X W^T + b).  X tiled [rows, K], rows = n_rows (multiple of 128 in storage).
int gemm_n(const dfb_model* m, const TiledBuf& X, long long n_rows, const PackedLayer& W, int relu, TiledBuf* out,
           float* out_f, long long ldo, int rows_per_patch, const float* bias_override, long long bias_patch_stride,
           cudaStream_t st, int epi = EPI_ACT, const TiledBuf* perpatch_B = nullptr, float* trans2 = nullptr) {
    GemmArgs g = base_args(m);
    g.A = X.p; g.a_plane = X.plane; g.KB = X.KB;
    if (perpatch_B) {
        g.B = perpatch_B->p; g.b_plane = perpatch_B->plane; g.b_patch_tiles = 1;
    } else {
        g.B = W.w.p; g.b_plane = W.w.plane;
    }
    g.tmode = 0;
    g.rows_per_patch = rows_per_patch;
    g.bias = bias_override ? bias_override : W.bias;
    g.bias_patch_stride = bias_patch_stride;
    g.relu = relu;
    g.m_valid = n_rows;
    const int n_out = perpatch_B ? 64 : W.out;
    g.n_valid = n_out;
    g.n_mma = n_out < 128 ? 64 : 128;
    if (out) { g.out_t = out->p; g.out_plane = out->plane; g.out_KB = out->KB; g.out_br = out->br; }
    g.out_f = out_f; g.ldo = ldo;
    g.trans2 = trans2;
    if (epi == EPI_DOT) { g.w4 = m->w_small[DFB_L_HEAD_CONV4]; }
    const int col_tiles = (n_out + 127) / 128;
    int NT = (epi == EPI_ACT && col_tiles % 2 == 0 && rows_per_patch != 0) ? 2 : 1;
    dim3 grid((unsigned)((n_rows + 127) / 128), (unsigned)(col_tiles / NT));
    return launch_gemm(g, NT, epi, grid, st);
}

template <int NT>
int launch_max_resident(GemmArgs g, long long n_patch, cudaStream_t st) {
    const int planes = fmt_planes(g.fmt);
    const size_t x_bytes = (size_t)planes * NT * g.KB * kBlkBytes;
    const size_t w_stage = (size_t)planes * kBlkBytes;
    const size_t budget = 227 * 1024 - 2048;
    int wstages = (int)((budget - x_bytes) / w_stage);
    if (wstages > 4) wstages = 4;
    const size_t smem = x_bytes + (size_t)wstages * w_stage;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_max_resident_kernel<NT>), (int)budget)) return rc;
    g.lbo = kLBO;
    g.sbo = kSBO;
    gemm_max_resident_kernel<NT><<<(unsigned)n_patch, kGemmThreads, smem, st>>>(g, wstages);
    note_launch();
    return check_cuda(cudaGetLastError(), "gemm_max_resident_kernel launch", __FILE__, __LINE__);
}

int g_head_fused = 1;  // debug switch (dfb_dbg_set key 2): 0 = separate 64->512 and 512->256 launches

static inline uint16_t host_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
    return (uint16_t)(u >> 16);
}
static inline float host_bf16_to_f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// HEAD_CONV1[:, 1024:1088] (512 x 64) -> 8 chunks of 64 output channels, each [hi 8 KB | lo 8 KB] in
// K-major core-matrix order [k-chunk j][row-group i][8 rows][8 elements] (LBO 1024 B, SBO 128 B).
int upload_head_w1_chunks(int fmt, const float* h_w, long long ldw, int col0, bf16** d_out, cudaStream_t st) {
    std::vector<uint16_t> buf((size_t)8 * 8192);
    unsigned char* bytes = reinterpret_cast<unsigned char*>(buf.data());
    for (int c = 0; c < 8; ++c)
        for (int r = 0; r < 64; ++r)
            for (int kc = 0; kc < 64; ++kc) {
                const float v = h_w[(size_t)(c * 64 + r) * ldw + col0 + kc];
                const size_t o = (size_t)c * 8192 + (size_t)(kc >> 3) * 512 + (size_t)(r >> 3) * 64 + (size_t)(r & 7) * 8 + (kc & 7);
                if (fmt == FMT_F16X3) {
                    const float vs = v * kWgtMain;
                    const __half h = __float2half_rn(vs);
                    const __half l = __float2half_rn(vs - __half2float(h));
                    buf[o] = *reinterpret_cast<const uint16_t*>(&h);
                    buf[o + 4096] = *reinterpret_cast<const uint16_t*>(&l);
                } else if (fmt != FMT_F16F8) {
                    const uint16_t hi = host_bf16(v);
                    const uint16_t lo = host_bf16(v - host_bf16_to_f(hi));
                    buf[o] = hi;
                    buf[o + 4096] = lo;
                } else {
                    // weight side of f16f8: main = fp16(w * 2^8); second plane per 16-kc group: chunk 2m = e4m3(wl * 2^15),
                    // chunk 2m + 1 = e4m3(wh * 2^3)
                    const __half h = __float2half_rn(v * kWgtMain);
                    buf[o] = *reinterpret_cast<const uint16_t*>(&h);
                    const float wh = __half2float(h) * (1.f / kWgtMain), wl = v - wh;
                    const size_t rowb = ((size_t)(r >> 3) * 64 + (size_t)(r & 7) * 8) * 2 + (size_t)(kc & 15);
                    const size_t base = ((size_t)c * 8192 + 4096) * 2;
                    const int m2 = kc >> 4;
                    bytes[base + (size_t)(2 * m2) * 1024 + rowb] = (unsigned char)__nv_cvt_float_to_fp8(wl * 32768.f, __NV_SATFINITE, __NV_E4M3);
                    bytes[base + (size_t)(2 * m2 + 1) * 1024 + rowb] = (unsigned char)__nv_cvt_float_to_fp8(wh * 8.f, __NV_SATFINITE, __NV_E4M3);
                }
            }
    DFB_CUDA(cudaMalloc(d_out, buf.size() * sizeof(uint16_t)));
    DFB_CUDA(cudaMemcpyAsync(*d_out, buf.data(), buf.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    DFB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int launch_head_fused(const dfb_model* m, const TiledBuf& pf, long long rows, const float* hg, const PackedLayer& W2,
                      TiledBuf& out, float* out_w, cudaStream_t st);

int g_chain_persistent = 1;  // dfb_dbg_set key 10: 0 = one CTA per patch
int g_chain = 1;  // debug switch (dfb_dbg_set key 5): 0 = separate pointwise / small-GEMM / max-pool launches

struct ChainSpec {
    int first_layer;            // DFB_L_* of the K=3 layer
    const float* R;
    float* pts_out;
    int n_layers;
    ChainLayer L[3];
    const PackedLayer* W3;
    int relu3;
};

ChainLayer chain_layer(const PackedLayer& pl, int relu) {
    ChainLayer c;
    memset(&c, 0, sizeof(c));
    c.w = pl.w.p; c.w_plane = pl.w.plane; c.bias = pl.bias; c.relu = relu; c.n_out = pl.out;
    c.h_bias = pl.h_bias.empty() ? nullptr : pl.h_bias.data();
    return c;
}

int launch_chain(const dfb_model* m, const float* patch, long long n_patch, const ChainSpec& sp, TiledBuf& gfeat, cudaStream_t st) {
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    a.patch = patch; a.R = sp.R; a.pts_out = sp.pts_out; a.k = m->k;
    memcpy(a.w0c, m->h_w_small[sp.first_layer].data(), sizeof(a.w0c));
    memcpy(a.b0c, m->L[sp.first_layer].h_bias.data(), sizeof(a.b0c));
    a.n_layers = sp.n_layers;
    for (int i = 0; i < sp.n_layers; ++i) {
        a.L[i] = sp.L[i];
        if (sp.L[i].h_bias) memcpy(a.biasc[i], sp.L[i].h_bias, sizeof(float) * sp.L[i].n_out);
    }
    a.w3 = sp.W3->w.p; a.w3_plane = sp.W3->w.plane; a.bias3 = sp.W3->bias; a.relu3 = sp.relu3; a.n_ch = sp.W3->out;
    a.out_t = gfeat.p; a.out_plane = gfeat.plane; a.out_KB = gfeat.KB;
    a.fmt = m->fmt; a.err = m->err;
    const int chain_id = sp.n_layers == 1 ? 0 : (sp.relu3 ? 1 : 2);   // QSTN, feature STN, encoder
    a.fmt_small = ((m->x1_mask >> (2 * chain_id)) & 1) ? FMT_F16X1 : m->fmt;
    a.fmt_max = ((m->x1_mask >> (2 * chain_id + 1)) & 1) ? FMT_F16X1 : m->fmt;
    a.tl = timeline_region(chain_id); a.tl_cta = g_tl_cta;
    // k = 512: a patch's 256 KB of activations do not fit shared memory, so it is processed as two 256-point work items
    const int ipp = m->k == 512 ? 2 : 1;
    const int NT = m->k / 128 / ipp;
    a.ipp = ipp;
    if (ipp > 1) {
        a.out_t = nullptr;
        a.out_f = m->chain_part;
    }
    n_patch *= ipp;
    const size_t smem = (size_t)2 * NT * 2 * kBlkBytes + 96 * 1024;
    if (int rc = ensure_dyn_smem(NT == 1 ? reinterpret_cast<const void*>(&chain_max_kernel<1>)
                                         : reinterpret_cast<const void*>(&chain_max_kernel<2>), (int)smem)) return rc;
    a.n_patch = (int)n_patch;
    // the per-tile barriers of the small layers complete n_layers + 1 times per patch: persistence needs that even
    const bool persistent = g_chain_persistent && (sp.n_layers & 1) && (sp.W3->out / 128) % 2 == 0;
    const unsigned grid = persistent ? std::min<unsigned>((unsigned)n_patch, (unsigned)sm_count()) : (unsigned)n_patch;
    if (NT == 1) chain_max_kernel<1><<<grid, kChainThreads, smem, st>>>(a);
    else chain_max_kernel<2><<<grid, kChainThreads, smem, st>>>(a);
    note_launch();
    if (ipp > 1) {
        const long long total = (n_patch / ipp) * a.n_ch;
        chain_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m->chain_part, n_patch / ipp, a.n_ch, ipp, gfeat.p, gfeat.plane,
                                                                              gfeat.KB, m->fmt);
        note_launch();
    }
    return check_cuda(cudaGetLastError(), "chain_max_kernel launch", __FILE__, __LINE__);
}

int g_max_resident = 1;  // debug switch (dfb_dbg_set key 1): 0 = one CTA per (patch, channel tile)

// T orientation with fused max-pool: g[p, c] = act(max_j (W X_p^T)[c, j] + b[c])
int gemm_t_max(const dfb_model* m, const PackedLayer& W, const TiledBuf& X, long long n_patch, int relu, TiledBuf* out,
               float* out_f, cudaStream_t st) {
    GemmArgs g = base_args(m);
    g.A = W.w.p; g.a_plane = W.w.plane; g.KB = W.w.KB;
    g.B = X.p; g.b_plane = X.plane;
    g.tmode = 1;
    g.bias = W.bias;
    g.relu = relu;
    g.m_valid = n_patch;
    g.n_valid = W.out;
    if (out) { g.out_t = out->p; g.out_plane = out->plane; g.out_KB = out->KB; g.out_br = out->br; }
    g.out_f = out_f; g.ldo = W.out;
    const int NT = m->k / 128;
    const size_t x_bytes = (size_t)fmt_planes(m->fmt) * NT * g.KB * kBlkBytes;
    if (g_max_resident && NT <= 2 && X.br == NT * 128 && x_bytes + 2 * (size_t)fmt_planes(m->fmt) * kBlkBytes <= 225 * 1024) {
        return NT == 1 ? launch_max_resident<1>(g, n_patch, st) : launch_max_resident<2>(g, n_patch, st);
    }
    if (X.br != 128) {
        set_error("gemm_t_max: the non-resident kernel needs 128-row activation blocks");
        return DFB_E_INTERNAL;
    }
    dim3 grid((unsigned)(W.out / 128), (unsigned)n_patch);
    return launch_gemm(g, NT, EPI_MAX, grid, st);
}

}  // namespace
}  // namespace dfb

namespace dfb {
int model_k(const dfb_model* m) { return m->k; }
}  // namespace dfb

extern "C" int dfb_model_profile(dfb_model* m, int32_t enable) {
    using namespace dfb;
    DFB_REQUIRE(m != nullptr, DFB_E_ARG, "dfb_model_profile: NULL model");
    m->prof = enable != 0;
    if (m->prof) {
        while (m->pool.size() < 8192) {  // created up front so the timed region only records
            cudaEvent_t e;
            DFB_CUDA(cudaEventCreate(&e));
            m->pool.push_back(e);
        }
    }
    return DFB_OK;
}

extern "C" int dfb_model_profile_read(dfb_model* m, double* ms, int64_t* launches, int32_t n_classes, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(m && ms && launches && n_classes >= DFB_PROF_NUM, DFB_E_ARG, "dfb_model_profile_read: bad arguments");
    DFB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    for (auto& r : m->recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
            ms[r.cls] += t;
            launches[r.cls] += 1;
        }
        m->pool.push_back(r.e0);
        m->pool.push_back(r.e1);
    }
    m->recs.clear();
    return DFB_OK;
}

// Debug / A-B switches (declared in include/deepfit_b200_debug.h, not in the ABI header): 0 swap LBO/SBO (descriptor probe), 1 resident max-pool kernel,
// 2 fused head, 4 layer 3 fused into the head, 5 fused chain kernels, 6 head experiment bits, 7 CTA / tile / patch whose
// timeline is recorded (-1 off), 8 persistent head, 9 per-tile entry/exit trace of the head, 10 persistent chain.
extern "C" DFB_API int dfb_dbg_set(int key, int value) {
    if (key == 0) dfb::g_swap_lbo_sbo = value;
    if (key == 1) dfb::g_max_resident = value;
    if (key == 2) dfb::g_head_fused = value;
    // key 3 (cluster-of-2 weight multicast in the head kernel) was removed with that variant: it lost its A/B
    if (key == 4) dfb::g_head_l3 = value;
    if (key == 5) dfb::g_chain = value;
    if (key == 6) dfb::g_head_dbg = value;
    if (key == 7) dfb::g_tl_cta = value;
    if (key == 8) dfb::g_head_persistent = value;
    if (key == 9) dfb::g_cta_trace_on = value;
    if (key == 10) dfb::g_chain_persistent = value;
    if (key == 11) dfb::g_gemm_cf = value;
    return DFB_OK;
}

// Profiling aid: (SM id, entry ns, exit ns) of the first n CTAs of the last traced head launch.
extern "C" DFB_API int dfb_dbg_cta_trace(long long* host_out, int n_cta) {
    using namespace dfb;
    DFB_REQUIRE(host_out && n_cta > 0 && n_cta <= 16384, DFB_E_ARG, "dfb_dbg_cta_trace: bad arguments");
    DFB_CUDA(cudaDeviceSynchronize());
    DFB_CUDA(cudaMemcpyFromSymbol(host_out, g_cta_trace, sizeof(long long) * 3 * n_cta));
    return DFB_OK;
}

// Profiling aid: copies the 4 x 64 clock64() stamps (chains 0-2, head) of the traced CTA to the host.
extern "C" DFB_API int dfb_dbg_timeline(long long* host_out, int n) {
    using namespace dfb;
    DFB_REQUIRE(host_out && n >= 4 * 64, DFB_E_ARG, "dfb_dbg_timeline: need room for 256 stamps");
    DFB_CUDA(cudaDeviceSynchronize());
    DFB_CUDA(cudaMemcpyFromSymbol(host_out, g_tl_dev, sizeof(long long) * 4 * 64));
    return DFB_OK;
}

extern "C" int dfb_model_destroy(dfb_model* m) {
    if (!m) return DFB_OK;
    for (void* p : m->allocs) cudaFree(p);
    for (auto& r : m->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : m->pool) cudaEventDestroy(e);
    delete m;
    return DFB_OK;
}

extern "C" int dfb_model_create(const dfb_model_desc* desc, dfb_stream stream, dfb_model** out) {
    using namespace dfb;
    DFB_REQUIRE(desc && out, DFB_E_ARG, "dfb_model_create: NULL argument");
    *out = nullptr;
    DFB_REQUIRE(desc->k >= 128 && desc->k <= 512 && desc->k % 128 == 0, DFB_E_UNSUPPORTED,
                "dfb_model_create: k=%d; the tensor-core path needs k in {128,256,384,512}", desc->k);
    DFB_REQUIRE(desc->weight_mode == DFB_WEIGHT_MODE_SIGMOID || desc->weight_mode == DFB_WEIGHT_MODE_TANH, DFB_E_UNSUPPORTED,
                "dfb_model_create: unsupported weight_mode %d", desc->weight_mode);
    static const int expect[DFB_NUM_LAYERS][2] = {{3, 64},     {64, 128},  {128, 1024}, {1024, 512}, {512, 256}, {256, 4},
                                                  {3, 64},     {64, 64},   {64, 64},    {64, 128},   {128, 1024}, {1024, 512},
                                                  {512, 256},  {256, 4096}, {64, 128},  {128, 1024}, {1088, 512}, {512, 256},
                                                  {256, 128},  {128, 1}};
    for (int i = 0; i < DFB_NUM_LAYERS; ++i) {
        DFB_REQUIRE(desc->layers[i].h_weight && desc->layers[i].h_bias, DFB_E_ARG, "dfb_model_create: layer %d has NULL data", i);
        DFB_REQUIRE(desc->layers[i].in_features == expect[i][0] && desc->layers[i].out_features == expect[i][1], DFB_E_ARG,
                    "dfb_model_create: layer %d is %dx%d, expected %dx%d", i, desc->layers[i].out_features,
                    desc->layers[i].in_features, expect[i][1], expect[i][0]);
    }
    int rc = dfb_device_check();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    dfb_model* m = new dfb_model();
    m->k = desc->k;
    m->weight_mode = desc->weight_mode;
    const int precision = desc->precision & 0xff, mix_mask = (desc->precision >> 8) & DFB_MIX_ALL;
    DFB_REQUIRE(precision >= DFB_PRECISION_BF16X3 && precision <= DFB_PRECISION_F16MIX && (desc->precision >> 8) == mix_mask &&
                    (mix_mask == 0 || precision == DFB_PRECISION_F16MIX), DFB_E_ARG,
                "dfb_model_create: unknown precision %d", desc->precision);
    static const int fmt_of[5] = {FMT_BF16X3, FMT_BF16, FMT_F16F8, FMT_F16X3, FMT_F16X3};
    m->fmt = fmt_of[precision];
    if (precision == DFB_PRECISION_F16MIX) m->x1_mask = mix_mask ? mix_mask : DFB_MIX_DEFAULT;
    m->max_batch = desc->max_batch > 0 ? (desc->max_batch + 127) / 128 * 128 : 1024;
#define DFB_TRY(expr)                     \
    do {                                  \
        rc = (expr);                      \
        if (rc) {                         \
            dfb_model_destroy(m);         \
            return rc;                    \
        }                                 \
    } while (0)
    auto track = [&](void* p) { m->allocs.push_back(p); };
    if (fmt_is_f16(m->fmt)) {
        // the main plane stores w * 2^8 in FP16: |w| must stay below 65504 / 256
        for (int i = 0; i < DFB_NUM_LAYERS; ++i) {
            const dfb_layer& l = desc->layers[i];
            float mx = 0.f;
            for (size_t j = 0; j < (size_t)l.in_features * l.out_features; ++j) mx = std::max(mx, fabsf(l.h_weight[j]));
            if (!(mx < 250.f)) {
                set_error("dfb_model_create: layer %d has |w| up to %g, outside the FP16 range of precisions f16x3 / f16f8; use bf16x3", i, mx);
                dfb_model_destroy(m);
                return DFB_E_UNSUPPORTED;
            }
        }
    }
    for (int i = 0; i < DFB_NUM_LAYERS; ++i) {
        const dfb_layer& l = desc->layers[i];
        DFB_TRY(upload_f32(l.h_bias, l.out_features, &m->L[i].bias, st));
        track(m->L[i].bias);
        m->L[i].in = l.in_features;
        m->L[i].out = l.out_features;
        m->L[i].h_bias.assign(l.h_bias, l.h_bias + l.out_features);
        if (i == DFB_L_QSTN_CONV1 || i == DFB_L_PF_CONV1) m->h_w_small[i].assign(l.h_weight, l.h_weight + 3 * 64);
        if (i == DFB_L_HEAD_CONV4) m->h_w_small[i].assign(l.h_weight, l.h_weight + 128);
        const bool small = (i == DFB_L_QSTN_CONV1 || i == DFB_L_PF_CONV1 || i == DFB_L_QSTN_FC3 || i == DFB_L_HEAD_CONV4);
        if (i == DFB_L_HEAD_CONV4) m->b4 = l.h_bias[0];
        if (small) {
            DFB_TRY(upload_f32(l.h_weight, (size_t)l.in_features * l.out_features, &m->w_small[i], st));
            track(m->w_small[i]);
        } else if (i == DFB_L_HEAD_CONV1) {
            DFB_TRY(upload_pack(m->fmt, l.h_weight, 512, 1088, 0, 1024, 1088, m->head_global, st));
            track(m->head_global.w.p);
            DFB_TRY(upload_pack(m->fmt, l.h_weight, 512, 1088, 1024, 64, 1088, m->head_point, st));
            track(m->head_point.w.p);
            m->head_global.bias = m->L[i].bias;
            m->head_point.bias = nullptr;
            DFB_TRY(upload_head_w1_chunks(m->fmt, l.h_weight, 1088, 1024, &m->head_w1c, st));
            track(m->head_w1c);
        } else if (i == DFB_L_STN_FC3) {
            // emit T2^T in the memory order of a tiled 64-row operand: output column
            // (kk>>3)*512 + (n>>3)*64 + (n&7)*8 + (kk&7) takes the reference's row kk*64+n; +I is symmetric
            std::vector<int> perm(4096);
            std::vector<float> hb(4096);
            for (int n = 0; n < 64; ++n)
                for (int kk = 0; kk < 64; ++kk) {
                    const int c = (kk >> 3) * 512 + (n >> 3) * 64 + (n & 7) * 8 + (kk & 7);
                    perm[c] = kk * 64 + n;
                    hb[c] = l.h_bias[kk * 64 + n] + (n == kk ? 1.f : 0.f);
                }
            DFB_TRY(upload_pack(m->fmt, l.h_weight, 4096, 256, 0, 256, 256, m->L[i], st, perm.data()));
            track(m->L[i].w.p);
            DFB_TRY(check_cuda(cudaMemcpyAsync(m->L[i].bias, hb.data(), 4096 * sizeof(float), cudaMemcpyHostToDevice, st),
                               "stn fc3 bias", __FILE__, __LINE__));
            DFB_TRY(check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__));
        } else {
            if (i == DFB_L_HEAD_CONV2) {
                DFB_TRY(upload_pack(m->fmt, l.h_weight, l.out_features, l.in_features, 0, l.in_features, l.in_features, m->head_w2_256, st,
                                    nullptr, 256));
                track(m->head_w2_256.w.p);
                m->head_w2_256.bias = m->L[i].bias;
            }
            float* keep_bias = m->L[i].bias;
            DFB_TRY(upload_pack(m->fmt, l.h_weight, l.out_features, l.in_features, 0, l.in_features, l.in_features, m->L[i], st));
            m->L[i].bias = keep_bias;
            track(m->L[i].w.p);
        }
    }
    // workspace
    const long long R = (long long)m->max_batch * m->k;
    const long long Bp = m->max_batch;
    // only x*T2 (x64c: encoder chain -> head) is materialised by the fused kernels; the activation buffers of
    // the unfused A/B-test paths (x64a, x64b, x128, h512, h256: ~1 MB per patch) are allocated on first use
    DFB_TRY(alloc_tiled(m->x64c, R, 64));
    track(m->x64c.p);
    DFB_TRY(alloc_tiled(m->gfeat, Bp, 1024)); track(m->gfeat.p);
    DFB_TRY(alloc_tiled(m->f512, Bp, 512)); track(m->f512.p);
    DFB_TRY(alloc_tiled(m->f256, Bp, 256)); track(m->f256.p);
    DFB_TRY(alloc_tiled(m->t2blk, Bp * 128, 64)); track(m->t2blk.p);
    // padded rows of FC inputs and the unused half of each per-patch T2 block must be finite
    DFB_TRY(check_cuda(cudaMemsetAsync(m->gfeat.p, 0, m->gfeat.plane * 2 * sizeof(bf16), st), "memset", __FILE__, __LINE__));
    DFB_TRY(check_cuda(cudaMemsetAsync(m->f512.p, 0, m->f512.plane * 2 * sizeof(bf16), st), "memset", __FILE__, __LINE__));
    DFB_TRY(check_cuda(cudaMemsetAsync(m->f256.p, 0, m->f256.plane * 2 * sizeof(bf16), st), "memset", __FILE__, __LINE__));
    DFB_TRY(check_cuda(cudaMemsetAsync(m->t2blk.p, 0, m->t2blk.plane * 2 * sizeof(bf16), st), "memset", __FILE__, __LINE__));
    DFB_TRY(check_cuda(cudaMalloc(&m->f256_f, (size_t)Bp * 256 * sizeof(float)), "malloc", __FILE__, __LINE__)); track(m->f256_f);
    DFB_TRY(check_cuda(cudaMalloc(&m->hg, (size_t)Bp * 512 * sizeof(float)), "malloc", __FILE__, __LINE__)); track(m->hg);
    if (m->k == 512) {
        DFB_TRY(check_cuda(cudaMalloc(&m->chain_part, (size_t)Bp * 2 * 1024 * sizeof(float)), "malloc", __FILE__, __LINE__));
        track(m->chain_part);
    }
    DFB_TRY(check_cuda(cudaMalloc(&m->err, sizeof(int)), "malloc", __FILE__, __LINE__)); track(m->err);
    DFB_TRY(check_cuda(cudaMemsetAsync(m->err, 0, sizeof(int), st), "memset", __FILE__, __LINE__));
    DFB_TRY(check_cuda(cudaStreamSynchronize(st), "sync", __FILE__, __LINE__));
#undef DFB_TRY
    *out = m;
    return DFB_OK;
}

extern "C" int dfb_model_status(dfb_model* m, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(m != nullptr, DFB_E_ARG, "dfb_model_status: NULL model");
    int h = 0;
    DFB_CUDA(cudaMemcpyAsync(&h, m->err, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
    DFB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    if (h != 0) {
        // report once and re-arm: later forwards on this model start with a clean flag
        DFB_CUDA(cudaMemsetAsync(m->err, 0, sizeof(int), as_stream(stream)));
        DFB_CUDA(cudaStreamSynchronize(as_stream(stream)));
        if (h == 100)
            set_error("activation magnitude above %g inside the network: outside the FP16 range of precisions f16x3 / f16f8 "
                      "(use precision bf16x3 for this model); outputs since the last status check are invalid", (double)kActLimit);
        else
            set_error("tensor-core pipeline time-out (mbarrier wait, role %d): the outputs of the forwards since the last "
                      "status check are invalid", h);
        return DFB_E_INTERNAL;
    }
    return DFB_OK;
}

extern "C" int dfb_model_forward(dfb_model* m, const float* d_patch, int64_t n, float* d_weights, float* d_trans,
                                 float* d_trans2, float* d_points_rot, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(m != nullptr, DFB_E_ARG, "dfb_model_forward: NULL model");
    DFB_REQUIRE(n >= 0, DFB_E_ARG, "dfb_model_forward: negative n");
    if (n == 0) return DFB_OK;
    DFB_REQUIRE(d_patch && d_weights && d_trans, DFB_E_ARG, "dfb_model_forward: NULL pointer argument");
    cudaStream_t st = as_stream(stream);
    const int k = m->k;
    const int split = m->fmt;   // operand format of the unfused pointwise kernels
    int rc;
#define DFB_RUN(expr)       \
    do {                    \
        rc = (expr);        \
        if (rc) return rc;  \
    } while (0)
    for (int64_t s = 0; s < n; s += m->max_batch) {
        const long long B = (n - s < m->max_batch) ? (n - s) : m->max_batch;
        const long long rows = B * k;
        const float* patch = d_patch + s * 3LL * k;
        float* trans = d_trans + s * 9;
        const unsigned pw_blocks = (unsigned)((rows + 127) / 128);

        const bool use_chain = g_chain && (k == 128 || k == 256 || k == 512);
        const bool head_all = g_head_fused && g_head_l3;
        if (!use_chain || !head_all) DFB_RUN(ensure_unfused_buffers(m));
        if (use_chain) {
            // ---- QSTN (DeepFit.py:410-441): points -> 64 -> 128 -> 1024 + max in one kernel ----
            {
                ProfScope ps(m, DFB_PROF_GEMM_MAX, st);
                ChainSpec sp;
                memset(&sp, 0, sizeof(sp));
                sp.first_layer = DFB_L_QSTN_CONV1;
                sp.n_layers = 1;
                sp.L[0] = chain_layer(m->L[DFB_L_QSTN_CONV2], 1);
                sp.W3 = &m->L[DFB_L_QSTN_CONV3];
                sp.relu3 = 1;
                DFB_RUN(launch_chain(m, patch, B, sp, m->gfeat, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->gfeat, B, m->L[DFB_L_QSTN_FC1], 1, &m->f512, nullptr, 0, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->f512, B, m->L[DFB_L_QSTN_FC2], 1, nullptr, m->f256_f, 256, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_QUAT, st);
                quat_kernel<<<(unsigned)((B + 3) / 4), 128, 0, st>>>(m->f256_f, B, m->w_small[DFB_L_QSTN_FC3], m->L[DFB_L_QSTN_FC3].bias, trans);
                note_launch();
            }
            // ---- feature STN (DeepFit.py:188-197, 353-380): rotate -> 64 -> 64 -> 64 -> 128 -> 1024 + max ----
            {
                ProfScope ps(m, DFB_PROF_GEMM_MAX, st);
                ChainSpec sp;
                memset(&sp, 0, sizeof(sp));
                sp.first_layer = DFB_L_PF_CONV1;
                sp.R = trans;
                sp.n_layers = 3;
                sp.L[0] = chain_layer(m->L[DFB_L_PF_CONV2], 1);
                sp.L[1] = chain_layer(m->L[DFB_L_STN_CONV1], 1);
                sp.L[2] = chain_layer(m->L[DFB_L_STN_CONV2], 1);
                sp.W3 = &m->L[DFB_L_STN_CONV3];
                sp.relu3 = 1;
                DFB_RUN(launch_chain(m, patch, B, sp, m->gfeat, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->gfeat, B, m->L[DFB_L_STN_FC1], 1, &m->f512, nullptr, 0, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->f512, B, m->L[DFB_L_STN_FC2], 1, &m->f256, nullptr, 0, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->f256, B, m->L[DFB_L_STN_FC3], 0, &m->t2blk, nullptr, 0, 0, nullptr, 0, st, EPI_T2, nullptr,
                               d_trans2 ? d_trans2 + s * 4096 : nullptr));
            }
            // ---- encoder (DeepFit.py:202-204, 233-235): rotate -> 64 -> 64 -> x*T2 (kept for the head) -> 128 -> 1024 + max ----
            {
                ProfScope ps(m, DFB_PROF_GEMM_MAX, st);
                ChainSpec sp;
                memset(&sp, 0, sizeof(sp));
                sp.first_layer = DFB_L_PF_CONV1;
                sp.R = trans;
                sp.pts_out = d_points_rot ? d_points_rot + s * 3LL * k : nullptr;
                sp.n_layers = 3;
                sp.L[0] = chain_layer(m->L[DFB_L_PF_CONV2], 1);
                memset(&sp.L[1], 0, sizeof(ChainLayer));
                sp.L[1].w = m->t2blk.p; sp.L[1].w_plane = m->t2blk.plane; sp.L[1].w_patch_stride = kBlkElems;
                sp.L[1].n_out = 64;
                sp.L[1].store_t = m->x64c.p; sp.L[1].store_plane = m->x64c.plane;
                sp.L[2] = chain_layer(m->L[DFB_L_ENC_CONV2], 1);
                sp.W3 = &m->L[DFB_L_ENC_CONV3];
                sp.relu3 = 0;
                DFB_RUN(launch_chain(m, patch, B, sp, m->gfeat, st));
            }
        } else {
            // ---- QSTN (DeepFit.py:410-441) ----
            {
                ProfScope ps(m, DFB_PROF_POINTWISE, st);
                pointwise3_kernel<<<pw_blocks, 128, 0, st>>>(patch, nullptr, B, k, m->w_small[DFB_L_QSTN_CONV1], m->L[DFB_L_QSTN_CONV1].bias,
                                                             m->x64a.p, m->x64a.plane, split, nullptr);
                note_launch();
            }
            {
                ProfScope ps(m, DFB_PROF_GEMM_SMALL, st);
                DFB_RUN(gemm_n(m, m->x64a, rows, m->L[DFB_L_QSTN_CONV2], 1, &m->x128, nullptr, 0, k, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_GEMM_MAX, st);
                DFB_RUN(gemm_t_max(m, m->L[DFB_L_QSTN_CONV3], m->x128, B, 1, &m->gfeat, nullptr, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->gfeat, B, m->L[DFB_L_QSTN_FC1], 1, &m->f512, nullptr, 0, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->f512, B, m->L[DFB_L_QSTN_FC2], 1, nullptr, m->f256_f, 256, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_QUAT, st);
                quat_kernel<<<(unsigned)((B + 3) / 4), 128, 0, st>>>(m->f256_f, B, m->w_small[DFB_L_QSTN_FC3], m->L[DFB_L_QSTN_FC3].bias, trans);
                note_launch();
            }

            // ---- point features (DeepFit.py:188-197) ----
            {
                ProfScope ps(m, DFB_PROF_POINTWISE, st);
                pointwise3_kernel<<<pw_blocks, 128, 0, st>>>(patch, trans, B, k, m->w_small[DFB_L_PF_CONV1], m->L[DFB_L_PF_CONV1].bias,
                                                             m->x64a.p, m->x64a.plane, split,
                                                             d_points_rot ? d_points_rot + s * 3LL * k : nullptr);
                note_launch();
            }
            {
                ProfScope ps(m, DFB_PROF_GEMM_SMALL, st);
                DFB_RUN(gemm_n(m, m->x64a, rows, m->L[DFB_L_PF_CONV2], 1, &m->x64b, nullptr, 0, k, nullptr, 0, st));
            }

            // ---- feature STN (DeepFit.py:353-380) ----
            {
                ProfScope ps(m, DFB_PROF_GEMM_SMALL, st);
                DFB_RUN(gemm_n(m, m->x64b, rows, m->L[DFB_L_STN_CONV1], 1, &m->x64c, nullptr, 0, k, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_GEMM_SMALL, st);
                DFB_RUN(gemm_n(m, m->x64c, rows, m->L[DFB_L_STN_CONV2], 1, &m->x128, nullptr, 0, k, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_GEMM_MAX, st);
                DFB_RUN(gemm_t_max(m, m->L[DFB_L_STN_CONV3], m->x128, B, 1, &m->gfeat, nullptr, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->gfeat, B, m->L[DFB_L_STN_FC1], 1, &m->f512, nullptr, 0, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->f512, B, m->L[DFB_L_STN_FC2], 1, &m->f256, nullptr, 0, 0, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_FC, st);
                DFB_RUN(gemm_n(m, m->f256, B, m->L[DFB_L_STN_FC3], 0, &m->t2blk, nullptr, 0, 0, nullptr, 0, st, EPI_T2, nullptr,
                               d_trans2 ? d_trans2 + s * 4096 : nullptr));
            }
            // x' = x T2 (DeepFit.py:202-204): per-patch B operand
            {
                ProfScope ps(m, DFB_PROF_GEMM_SMALL, st);
                DFB_RUN(gemm_n(m, m->x64b, rows, m->head_point, 0, &m->x64c, nullptr, 0, k, nullptr, 0, st, EPI_ACT, &m->t2blk));
            }

            // ---- encoder (DeepFit.py:233-235) ----
            {
                ProfScope ps(m, DFB_PROF_GEMM_SMALL, st);
                DFB_RUN(gemm_n(m, m->x64c, rows, m->L[DFB_L_ENC_CONV2], 1, &m->x128, nullptr, 0, k, nullptr, 0, st));
            }
            {
                ProfScope ps(m, DFB_PROF_GEMM_MAX, st);
                DFB_RUN(gemm_t_max(m, m->L[DFB_L_ENC_CONV3], m->x128, B, 0, &m->gfeat, nullptr, st));
            }

        }

        // ---- head (DeepFit.py:304-316): conv1 split into global (per patch) + point (per point) parts ----
        {
            ProfScope ps(m, DFB_PROF_FC, st);
            DFB_RUN(gemm_n(m, m->gfeat, B, m->head_global, 0, nullptr, m->hg, 512, 0, nullptr, 0, st));
        }
        if (g_head_fused) {
            ProfScope ps(m, DFB_PROF_HEAD, st);
            DFB_RUN(launch_head_fused(m, m->x64c, rows, m->hg, m->head_w2_256, m->h256,
                                      head_all ? d_weights + s * (long long)k : nullptr, st));
        } else {
            {
                ProfScope ps(m, DFB_PROF_HEAD, st);
                DFB_RUN(gemm_n(m, m->x64c, rows, m->head_point, 1, &m->h512, nullptr, 0, k, m->hg, 512, st));
            }
            {
                ProfScope ps(m, DFB_PROF_HEAD, st);
                DFB_RUN(gemm_n(m, m->h512, rows, m->L[DFB_L_HEAD_CONV2], 1, &m->h256, nullptr, 0, k, nullptr, 0, st));
            }
        }
        if (!head_all) {
            ProfScope ps(m, DFB_PROF_HEAD, st);
            DFB_RUN(gemm_n(m, m->h256, rows, m->L[DFB_L_HEAD_CONV3], 1, nullptr, d_weights + s * (long long)k, 0, k, nullptr, 0, st,
                           EPI_DOT));
        }
        DFB_RUN(check_cuda(cudaGetLastError(), "model forward launches", __FILE__, __LINE__));
    }
#undef DFB_RUN
    return DFB_OK;
}

// Test hook: unpack one workspace buffer of the last forward to fp32 row-major [rows, cols].
// which: 0 x64a (pf conv1), 1 x64b (pf conv2), 2 x64c (x*T2), 3 x128 (enc conv2), 4 h512, 5 h256,
//        6 gfeat, 7 f512, 8 f256
extern "C" DFB_API int dfb_dbg_model_dump(dfb_model* m, int32_t which, int64_t rows, int32_t cols, float* d_out, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(m && d_out, DFB_E_ARG, "dfb_dbg_model_dump: NULL argument");
    TiledBuf* bufs[] = {&m->x64a, &m->x64b, &m->x64c, &m->x128, &m->h512, &m->h256, &m->gfeat, &m->f512, &m->f256};
    DFB_REQUIRE(which >= 0 && which < 9, DFB_E_ARG, "dfb_dbg_model_dump: bad buffer id");
    TiledBuf* t = bufs[which];
    DFB_REQUIRE(t->p != nullptr, DFB_E_ARG, "dfb_dbg_model_dump: buffer %d is not materialised in this configuration", which);
    DFB_REQUIRE(rows <= t->rows_pad && cols <= t->KB * 64, DFB_E_ARG, "dfb_dbg_model_dump: shape exceeds the buffer");
    const long long total = rows * (long long)cols;
    unpack_tiled_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(t->p, t->plane, m->fmt, rows, cols, t->KB, d_out, t->br);
    return check_cuda(cudaGetLastError(), "unpack", __FILE__, __LINE__);
}

// Test hook: C[M,N] = A[M,K] B[N,K]^T (+bias, relu) through the tcgen05 kernel, fp32 row-major in/out.
// mode 0: N orientation, EPI_ACT, out [M,N];  mode 1: T orientation + max over groups of k_pts rows of A... see tests.
extern "C" DFB_API int dfb_dbg_gemm(const float* d_A, const float* d_B, const float* d_bias, int64_t M, int32_t N, int32_t K,
                                    int32_t mode, int32_t nprod, int32_t relu, int32_t k_pts, float* d_out,
                                    float* d_out_tiled_unpacked, dfb_stream stream) {
    using namespace dfb;
    int rc = dfb_device_check();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    dfb_model fake;
    // nprod = tensor-core product units per k-step; 4 selects the split-FP16 form of the three-product scheme
    const int fmt = nprod == 1 ? FMT_BF16 : (nprod == 2 ? FMT_F16F8 : (nprod == 4 ? FMT_F16X3 : FMT_BF16X3));
    fake.fmt = fmt;
    fake.k = k_pts;
    DFB_CUDA(cudaMalloc(&fake.err, sizeof(int)));
    DFB_CUDA(cudaMemsetAsync(fake.err, 0, sizeof(int), st));
    TiledBuf tA, tB, tO;
    rc = alloc_tiled(tA, M, K, (mode == 1 && k_pts == 256) ? 256 : 128); if (rc) return rc;
    rc = alloc_tiled(tB, N, K); if (rc) return rc;
    auto pack = [&](const float* src, long long rows, int cols, TiledBuf& t, int wside) {
        const long long total = t.rows_pad * (t.KB * 8);
        pack_tiled_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, rows, cols, cols, t.rows_pad, t.KB * 64, t.p, t.plane, fmt, wside, t.br);
    };
    pack(d_A, M, K, tA, 0);   // activations
    pack(d_B, N, K, tB, 1);   // weights
    PackedLayer W;
    W.bias = const_cast<float*>(d_bias);
    if (mode == 0) {
        W.w = tB; W.in = K; W.out = N;
        rc = alloc_tiled(tO, M, N); if (rc) return rc;
        cudaMemsetAsync(tO.p, 0, tO.plane * 2 * sizeof(bf16), st);
        rc = gemm_n(&fake, tA, M, W, relu, &tO, d_out, N, k_pts, nullptr, 0, st);
        if (rc == 0 && d_out_tiled_unpacked) {
            const long long total = M * (long long)N;
            unpack_tiled_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tO.p, tO.plane, fmt, M, N, tO.KB, d_out_tiled_unpacked, tO.br);
        }
    } else {
        // A = activations [M = n_patch*k_pts, K], B = weights [N, K]; out [n_patch, N]
        W.w = tB; W.in = K; W.out = N;
        const long long n_patch = M / k_pts;
        rc = alloc_tiled(tO, n_patch, N); if (rc) return rc;
        cudaMemsetAsync(tO.p, 0, tO.plane * 2 * sizeof(bf16), st);
        rc = gemm_t_max(&fake, W, tA, n_patch, relu, &tO, d_out, st);
        if (rc == 0 && d_out_tiled_unpacked) {
            const long long total = n_patch * (long long)N;
            unpack_tiled_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tO.p, tO.plane, fmt, n_patch, N, tO.KB, d_out_tiled_unpacked, tO.br);
        }
    }
    int h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, fake.err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tA.p); cudaFree(tB.p); cudaFree(tO.p); cudaFree(fake.err);
    if (e != cudaSuccess) return check_cuda(e, "dfb_dbg_gemm", __FILE__, __LINE__);
    if (rc) return rc;
    if (h != 0) {
        set_error("dfb_dbg_gemm: pipeline time-out, role %d", h);
        return DFB_E_INTERNAL;
    }
    return DFB_OK;
}
