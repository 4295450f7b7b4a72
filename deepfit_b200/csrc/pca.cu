// Stage 2: patch gather, centring, scaling and PCA alignment -- one warp per patch.
//
// Replaces SinglePointCloudDataset.__getitem__ lines 23-30 and pca_points (reference
// tutorial/tutorial_utils.py:23-30,35-57; dataset.py:320-366):
//     P = (points[idx] - points[centre]) / rad      fp32, IEEE division by float(rad)
//     m = mean(P);  U = eigenvectors of (P-m)^T (P-m) by descending eigenvalue
//     out = P U  (the reference evaluates (P-m) U + m U), written planar [3,k]
// The covariance and the 3x3 symmetric eigen-solve (cyclic Jacobi) run in fp64 registers, every lane
// redundantly, so U is the exact-arithmetic answer to ~1e-15 -- LAPACK's own fp32 SVD is farther from
// it than we are.  Column signs are library-defined in the reference; here
// each axis is signed so that its largest-magnitude component is positive.
#include "common.cuh"

namespace dfb {
namespace {

constexpr int kPcaWarps = 8;
constexpr int kMaxPerLane = 32;  // k <= 1024

struct PcaArgs {
    const float* pts;
    const int* idx;
    const int* query;
    long long q_begin, q_count;
    int k;
    const double* rad;
    int scale, do_pca;
    float* patch;
    float* U;
    const int* valid;   // NULL, or per patch the number of leading real neighbours: the rest are padding (zero rows)
    const float4* pts4; // F4 kernels: the cloud as float4 (x, y, z, -) in original order, one 16-byte load per neighbour
    double* cov;        // SPLIT kernels: [q_count,6] upper triangle of the patch covariance (c00 c01 c02 c11 c12 c22)
};

__device__ __forceinline__ void jacobi_rotate(double (&A)[3][3], double (&V)[3][3], int p, int q) {
    const double apq = A[p][q];
    if (fabs(apq) < 1e-300) return;
    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    const int r = 3 - p - q;
    const double app = A[p][p], aqq = A[q][q], arp = A[r][p], arq = A[r][q];
    A[p][p] = app - t * apq;
    A[q][q] = aqq + t * apq;
    A[p][q] = A[q][p] = 0.0;
    A[r][p] = A[p][r] = c * arp - s * arq;
    A[r][q] = A[q][r] = s * arp + c * arq;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double vip = V[i][p], viq = V[i][q];
        V[i][p] = c * vip - s * viq;
        V[i][q] = s * vip + c * viq;
    }
}

// Eigenvectors of the symmetric 3x3 covariance by descending eigenvalue (cyclic Jacobi in fp64), each signed so that its
// largest-magnitude component is positive, rounded to fp32: U[r][c] = component r of axis c.  ONE definition for the
// fused kernel (every lane of a warp, redundantly) and for the split path's one-thread-per-patch kernel, so both give
// the same bits.
__device__ __forceinline__ void eig_frame(double c00, double c01, double c02, double c11, double c12, double c22, float (&Uf)[3][3]) {
    double A[3][3] = {{c00, c01, c02}, {c01, c11, c12}, {c02, c12, c22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double dia = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-18 * dia) break;
        jacobi_rotate(A, V, 0, 1);
        jacobi_rotate(A, V, 0, 2);
        jacobi_rotate(A, V, 1, 2);
    }
    // order columns by descending eigenvalue (stable 3-element network, register-only)
    double e0 = A[0][0], e1 = A[1][1], e2 = A[2][2];
#define DFB_SWAP_COLS(ea, eb, ca, cb)                                   \
    if (ea < eb) {                                                  \
        double tmp = ea; ea = eb; eb = tmp;                         \
        _Pragma("unroll") for (int i = 0; i < 3; ++i) {             \
            tmp = V[i][ca]; V[i][ca] = V[i][cb]; V[i][cb] = tmp;    \
        }                                                           \
    }
    DFB_SWAP_COLS(e0, e1, 0, 1)
    DFB_SWAP_COLS(e1, e2, 1, 2)
    DFB_SWAP_COLS(e0, e1, 0, 1)
#undef DFB_SWAP_COLS
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double v0 = V[0][c], v1 = V[1][c], v2 = V[2][c];
        // sign: largest-magnitude component positive (first one on ties)
        double big = v0;
        if (fabs(v1) > fabs(big)) big = v1;
        if (fabs(v2) > fabs(big)) big = v2;
        const double sgn = big < 0.0 ? -1.0 : 1.0;
        Uf[0][c] = (float)(sgn * v0); Uf[1][c] = (float)(sgn * v1); Uf[2][c] = (float)(sgn * v2);
    }
}

// NPL: points per lane (compile-time upper bound, k <= 32*NPL).  F4: gather from the float4 copy of the cloud (one
// 16-byte load per neighbour instead of three 4-byte ones).  SPLIT: stop after the covariance -- the patch is written
// centred and scaled but NOT rotated, the covariance goes to a.cov; pca_eig_kernel turns it into U with one thread per
// patch (the Jacobi sweeps are a serial fp64 chain that the fused kernel runs 32 times redundantly per patch), and the
// jet-fit kernel applies U while it reads the patch (it has a per-patch rotation input for the QSTN already).
template <int NPL, bool F4, bool SPLIT>
__global__ void __launch_bounds__(kPcaWarps * 32) pca_kernel(PcaArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = a.k;
    for (long long qw = (long long)blockIdx.x * kPcaWarps + warp; qw < a.q_count; qw += (long long)gridDim.x * kPcaWarps) {
        const long long ci = a.query ? (long long)a.query[qw] : a.q_begin + qw;
        float cx, cy, cz;
        if (F4) {
            const float4 c4 = __ldg(a.pts4 + ci);
            cx = c4.x; cy = c4.y; cz = c4.z;
        } else {
            cx = a.pts[ci * 3]; cy = a.pts[ci * 3 + 1]; cz = a.pts[ci * 3 + 2];
        }
        const float r32 = a.scale ? (float)a.rad[qw] : 1.f;
        const int* row = a.idx + qw * (long long)k;
        const int nv = a.valid ? min(max(a.valid[qw], 0), k) : k;   // rows >= nv are zero padding, outside mean / covariance
        float X[NPL], Y[NPL], Z[NPL];
        double sx = 0.0, sy = 0.0, sz = 0.0;
        // Three passes so that every load of the patch is in flight at once: all neighbour indices (coalesced), then all
        // point gathers (L2 latency, paid once per patch instead of once per point -- written as one loop the compiler
        // serialised index load -> gather -> IEEE division, whose slow-path branch fences the schedule), then the arithmetic.
        int id[NPL];
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int j = lane + 32 * t;
            id[t] = j < nv ? __ldg(row + j) : -1;
        }
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            X[t] = Y[t] = Z[t] = 0.f;
            if (F4) {
                if (id[t] >= 0) {
                    const float4 p4 = __ldg(a.pts4 + id[t]);
                    X[t] = p4.x; Y[t] = p4.y; Z[t] = p4.z;
                }
            } else {
                if (id[t] >= 0) {
                    const float* pp = a.pts + (long long)id[t] * 3;
                    X[t] = __ldg(pp); Y[t] = __ldg(pp + 1); Z[t] = __ldg(pp + 2);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            if (id[t] >= 0) {
                float x = X[t] - cx, y = Y[t] - cy, z = Z[t] - cz;
                if (a.scale) {
                    x = __fdiv_rn(x, r32); y = __fdiv_rn(y, r32); z = __fdiv_rn(z, r32);
                }
                X[t] = x; Y[t] = y; Z[t] = z;
                sx += x; sy += y; sz += z;
            }
        }
        float* out = a.patch + qw * 3LL * k;
        if (!a.do_pca) {
#pragma unroll
            for (int t = 0; t < NPL; ++t) {
                const int j = lane + 32 * t;
                if (j < k) { out[j] = X[t]; out[k + j] = Y[t]; out[2 * k + j] = Z[t]; }
            }
            if (a.U && lane < 9) a.U[qw * 9 + lane] = (lane % 4 == 0) ? 1.f : 0.f;
            continue;
        }
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
        const double mx = sx / max(nv, 1), my = sy / max(nv, 1), mz = sz / max(nv, 1);
        double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            if (lane + 32 * t < nv) {
                const double dx = X[t] - mx, dy = Y[t] - my, dz = Z[t] - mz;
                c00 += dx * dx; c01 += dx * dy; c02 += dx * dz;
                c11 += dy * dy; c12 += dy * dz; c22 += dz * dz;
            }
        }
        c00 = warp_sum(c00); c01 = warp_sum(c01); c02 = warp_sum(c02);
        c11 = warp_sum(c11); c12 = warp_sum(c12); c22 = warp_sum(c22);
        if (SPLIT) {
#pragma unroll
            for (int t = 0; t < NPL; ++t) {
                const int j = lane + 32 * t;
                if (j < k) { out[j] = X[t]; out[k + j] = Y[t]; out[2 * k + j] = Z[t]; }
            }
            if (lane == 0) {
                double* cv = a.cov + qw * 6;
                cv[0] = c00; cv[1] = c01; cv[2] = c02; cv[3] = c11; cv[4] = c12; cv[5] = c22;
            }
            continue;
        }
        float Uf[3][3];
        eig_frame(c00, c01, c02, c11, c12, c22, Uf);
#pragma unroll
        for (int t = 0; t < NPL; ++t) {
            const int j = lane + 32 * t;
            if (j < k) {
                const float x = X[t], y = Y[t], z = Z[t];
                out[j] = fmaf(z, Uf[2][0], fmaf(y, Uf[1][0], x * Uf[0][0]));
                out[k + j] = fmaf(z, Uf[2][1], fmaf(y, Uf[1][1], x * Uf[0][1]));
                out[2 * k + j] = fmaf(z, Uf[2][2], fmaf(y, Uf[1][2], x * Uf[0][2]));
            }
        }
        if (a.U && lane == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) a.U[qw * 9 + r * 3 + c] = Uf[r][c];
        }
    }
}

// The split path's second step: one thread per patch, covariance -> U (row-major f32[9]).
__global__ void __launch_bounds__(128) pca_eig_kernel(const double* __restrict__ cov, long long n, float* __restrict__ U) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* cv = cov + i * 6;
    float Uf[3][3];
    eig_frame(cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], Uf);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) U[i * 9 + r * 3 + c] = Uf[r][c];
}

template <bool F4, bool SPLIT>
void launch_pca(const PcaArgs& a, unsigned blocks, cudaStream_t st) {
    const int npl = (a.k + 31) / 32;
    if (npl <= 4) pca_kernel<4, F4, SPLIT><<<blocks, kPcaWarps * 32, 0, st>>>(a);
    else if (npl <= 8) pca_kernel<8, F4, SPLIT><<<blocks, kPcaWarps * 32, 0, st>>>(a);
    else if (npl <= 16) pca_kernel<16, F4, SPLIT><<<blocks, kPcaWarps * 32, 0, st>>>(a);
    else pca_kernel<32, F4, SPLIT><<<blocks, kPcaWarps * 32, 0, st>>>(a);
}

}  // namespace

// d_points4: NULL, or the float4 copy of the cloud (grid_points4) -- same values, one load per neighbour.
// d_cov: NULL = the fused kernel (patch rotated into the PCA frame, U written); otherwise the split path: the patch is
// written centred / scaled but unrotated, the covariance goes to d_cov, and a second kernel turns it into d_U.
int patch_pca_impl(const float* d_points, const float4* d_points4, int64_t n_points, const int32_t* d_idx, const int32_t* d_valid,
                   const int32_t* d_query, int64_t q_begin, int64_t q_count, int32_t k, const double* d_rad, int32_t scale_by_rad,
                   float* d_patch, float* d_U, double* d_cov, dfb_stream stream) {
    DFB_REQUIRE(k >= 1 && k <= 32 * kMaxPerLane, DFB_E_ARG, "dfb_patch_pca: k=%d out of range (1..1024)", k);
    DFB_REQUIRE(q_count >= 0, DFB_E_ARG, "dfb_patch_pca: negative q_count");
    if (q_count == 0) return DFB_OK;
    DFB_REQUIRE(d_points && d_idx && d_patch, DFB_E_ARG, "dfb_patch_pca: NULL pointer argument");
    DFB_REQUIRE(!(scale_by_rad & 1) || d_rad != nullptr, DFB_E_ARG, "dfb_patch_pca: scale_by_rad set but d_rad is NULL");
    DFB_REQUIRE(d_query != nullptr || (q_begin >= 0 && q_begin + q_count <= n_points), DFB_E_ARG,
                "dfb_patch_pca: query range outside the cloud");
    DFB_REQUIRE(d_cov == nullptr || (d_U != nullptr && !(scale_by_rad & 2)), DFB_E_ARG, "dfb_patch_pca: the split path needs d_U and PCA");
    int rc = dfb_device_check();
    if (rc) return rc;
    PcaArgs a{d_points, d_idx, d_query, q_begin, q_count, k, d_rad, scale_by_rad & 1, (scale_by_rad & 2) ? 0 : 1, d_patch, d_U, d_valid,
              d_points4, d_cov};
    long long blocks = (q_count + kPcaWarps - 1) / kPcaWarps;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = as_stream(stream);
    if (d_cov) {
        if (d_points4) launch_pca<true, true>(a, (unsigned)blocks, st);
        else launch_pca<false, true>(a, (unsigned)blocks, st);
        pca_eig_kernel<<<(unsigned)((q_count + 127) / 128), 128, 0, st>>>(d_cov, (long long)q_count, d_U);
        note_launch(2);
    } else {
        if (d_points4) launch_pca<true, false>(a, (unsigned)blocks, st);
        else launch_pca<false, false>(a, (unsigned)blocks, st);
        note_launch();
    }
    return check_cuda(cudaGetLastError(), "pca_kernel launch", __FILE__, __LINE__);
}
}  // namespace dfb

extern "C" int dfb_patch_pca_ragged(const float* d_points, int64_t n_points, const int32_t* d_idx, const int32_t* d_valid,
                                    const int32_t* d_query, int64_t q_begin, int64_t q_count, int32_t k, const double* d_rad,
                                    int32_t scale_by_rad, float* d_patch, float* d_U, dfb_stream stream) {
    return dfb::patch_pca_impl(d_points, nullptr, n_points, d_idx, d_valid, d_query, q_begin, q_count, k, d_rad, scale_by_rad, d_patch,
                               d_U, nullptr, stream);
}

extern "C" int dfb_patch_pca(const float* d_points, int64_t n_points, const int32_t* d_idx, const int32_t* d_query,
                             int64_t q_begin, int64_t q_count, int32_t k, const double* d_rad, int32_t scale_by_rad,
                             float* d_patch, float* d_U, dfb_stream stream) {
    return dfb_patch_pca_ragged(d_points, n_points, d_idx, nullptr, d_query, q_begin, q_count, k, d_rad, scale_by_rad, d_patch, d_U,
                                stream);
}
