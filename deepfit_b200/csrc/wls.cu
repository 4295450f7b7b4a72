// Stage 4+5: weighted n-jet fit, normal and principal curvatures in one kernel.
//
// Replaces fit_Wjet / solve_linear_system (reference models/DeepFit.py:11-154), the normal at :88,
// compute_principal_curvatures (tutorial/tutorial_utils.py:196-226) and the caller epilogue
// (tutorial/compute_normals.py:67,79; test_n_est.py:154-161).
//
// Design (memory-bound stage, roofline = HBM): the normal matrix of a monomial basis is a moment
// matrix, so instead of materialising the Vandermonde matrix A [k,M] (10 KB/patch) and the 55/120
// upper-triangle entries of A^T W A, each point contributes to the (2n+1)(2n+2)/2 distinct moments
// sum w x^a y^b (a+b <= 2n) and the (n+1)(n+2)/2 moments sum w z x^a y^b (a+b <= n).
//   phase 1: 8 lanes per patch (4 patches per warp, 128-bit coalesced loads of the planar
//            x/y/z/w rows, the next chunk prefetched), register accumulators, 3 xor-shuffle steps,
//            totals parked in shared memory.  The moments are SPLIT BETWEEN THE TWO WARPS of a
//            CTA by the parity of the x power: all 38 (order 3) fp64 accumulators in one thread
//            cost 233 registers, i.e. 8 warps per SM and an FP64 pipe 34 % busy behind exposed
//            load latency; half of them (plus the powers that half needs, +9 % fp64 work) fit
//            ~100 registers.  Both warps read the same rows (the second read hits L1 / L2).
//   phase 2: one lane per patch (each warp solves half of the CTA's batch): CGAL-style
//            preconditioning applied to the moments, XtX/XtY
// Numerics: a moment matrix, unlike the Gram matrix A^T W A the reference forms, is not backward
// stable in fp32 -- entry-wise rounding acts on the solution with cond(XtX) (up to 4e5 at order 4 on
// the lidar tutorial cloud) instead of cond(A).  The moments are therefore accumulated in fp64
// (B200 runs DFMA at half the FFMA rate, so the stage stays HBM-bound) and the fp32 Cholesky factor is
// used as the preconditioner of an fp64-residual refinement; the result is the exact least-squares
// solution of the fp32 inputs to ~1e-9, i.e. at most the reference's own rounding error away from it.
#include "common.cuh"

#include <stdlib.h>

namespace dfb {
namespace {

constexpr int kWarpsPerBlock = 2;

template <int ORDER>
struct Jet {
    static constexpr int D = 2 * ORDER;
    static constexpr int NM = (D + 1) * (D + 2) / 2;          // moments of w x^a y^b
    static constexpr int NZ = (ORDER + 1) * (ORDER + 2) / 2;  // moments of w z x^a y^b
    static constexpr int M = NZ;                              // unknowns
    __host__ __device__ static constexpr int mi(int a, int b) { return a * (D + 1) - a * (a - 1) / 2 + b; }
    __host__ __device__ static constexpr int zi(int a, int b) { return a * (ORDER + 1) - a * (a - 1) / 2 + b; }
};
// doubles per patch slot: the moments + sum|x|, sum|y|, rounded up to an odd count (conflict-free for one lane per slot)
template <int ORDER>
constexpr int slot_stride() { return (Jet<ORDER>::NM + Jet<ORDER>::NZ + 2) | 1; }

// Column monomials (power of x, power of y) in the reference's column order, DeepFit.py:51-76.
template <int ORDER>
__device__ __forceinline__ constexpr int colx(int i) {
    constexpr int8_t t[4][15] = {{1, 0, 0}, {1, 0, 2, 0, 1, 0}, {1, 0, 2, 0, 1, 3, 0, 2, 1, 0},
                                 {1, 0, 2, 0, 1, 3, 0, 2, 1, 4, 0, 3, 1, 2, 0}};
    return t[ORDER - 1][i];
}
template <int ORDER>
__device__ __forceinline__ constexpr int coly(int i) {
    constexpr int8_t t[4][15] = {{0, 1, 0}, {0, 1, 0, 2, 1, 0}, {0, 1, 0, 2, 1, 0, 3, 1, 2, 0},
                                 {0, 1, 0, 2, 1, 0, 3, 1, 2, 0, 4, 1, 3, 2, 0}};
    return t[ORDER - 1][i];
}

// Which warp of the CTA accumulates what (PAR = warp index): the moments sum w x^a y^b go by the parity of a; of the
// sum w z x^a y^b the odd warp also takes a = 0, which evens out the two halves (order 3: 31 vs 28 fp64 operations per
// point where an unsplit thread does 54).
template <int PAR>
__host__ __device__ constexpr bool owns_m(int a) { return (a & 1) == PAR; }
template <int PAR>
__host__ __device__ constexpr bool owns_z(int a) { return PAR == 1 ? ((a & 1) == 1 || a == 0) : ((a & 1) == 0 && a >= 2); }

template <int ORDER>
struct Acc {
    double m[Jet<ORDER>::NM];   // only the entries this warp owns are ever touched (the others cost no register)
    double zm[Jet<ORDER>::NZ];
    float sax, say;  // sum |x|, sum |y|   (odd warp)
    float nvalid;    // number of weights > 1e-3
};

template <int ORDER, int PAR>
__device__ __forceinline__ void acc_point(Acc<ORDER>& A, float x, float y, float z, float w, float wraw) {
    using J = Jet<ORDER>;
    double xp[J::D + 1], yp[J::D + 1];
    const double xd = x, yd = y, zd = z;
    const double x2 = xd * xd;
    xp[0] = w;
    xp[1] = xp[0] * xd;
#pragma unroll
    for (int a = 2; a <= J::D; ++a) xp[a] = xp[a - 2] * x2;   // each parity chain on its own; unused powers are dead code
    yp[0] = 1.0;
    yp[1] = yd;
#pragma unroll
    for (int b = 2; b <= J::D; ++b) yp[b] = yp[b - 1] * yd;
#pragma unroll
    for (int a = 0; a <= J::D; ++a) {
        if (!owns_m<PAR>(a)) continue;
#pragma unroll
        for (int b = 0; b <= J::D - a; ++b) A.m[J::mi(a, b)] = b == 0 ? A.m[J::mi(a, b)] + xp[a] : fma(xp[a], yp[b], A.m[J::mi(a, b)]);
    }
#pragma unroll
    for (int a = 0; a <= ORDER; ++a) {
        if (!owns_z<PAR>(a)) continue;
        const double xz = xp[a] * zd;
#pragma unroll
        for (int b = 0; b <= ORDER - a; ++b) A.zm[J::zi(a, b)] = b == 0 ? A.zm[J::zi(a, b)] + xz : fma(xz, yp[b], A.zm[J::zi(a, b)]);
    }
    if (PAR == 1) {
        A.sax += fabsf(x);
        A.say += fabsf(y);
    }
    A.nvalid += (wraw > 1e-3f) ? 1.f : 0.f;
}

__device__ __forceinline__ void rotate(const float* R, float& x, float& y, float& z) {
    // row vector times matrix, reference DeepFit.py:188-190 (bmm(points^T, trans))
    const float xr = fmaf(z, R[6], fmaf(y, R[3], x * R[0]));
    const float yr = fmaf(z, R[7], fmaf(y, R[4], x * R[1]));
    const float zr = fmaf(z, R[8], fmaf(y, R[5], x * R[2]));
    x = xr; y = yr; z = zr;
}

// Accumulate one patch with the 8 lanes of a group.  sub = lane & 7.
template <int ORDER, int PAR>
__device__ __forceinline__ void acc_patch(Acc<ORDER>& A, const float* __restrict__ px, const float* __restrict__ pw,
                                          const float* R, int k, int sub, bool use_w) {
#pragma unroll
    for (int i = 0; i < Jet<ORDER>::NM; ++i) A.m[i] = 0.0;
#pragma unroll
    for (int i = 0; i < Jet<ORDER>::NZ; ++i) A.zm[i] = 0.0;
    A.sax = A.say = A.nvalid = 0.f;
    const float* py = px + k;
    const float* pz = px + 2 * k;
    if ((k & 3) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(px);
        const float4* y4 = reinterpret_cast<const float4*>(py);
        const float4* z4 = reinterpret_cast<const float4*>(pz);
        const float4* w4 = reinterpret_cast<const float4*>(pw);
        const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
        const int nc = k >> 2;
        // the next chunk's loads are in flight while this one is accumulated
        float4 Xn = one4, Yn = one4, Zn = one4, Wn = one4;
        if (sub < nc) {
            Xn = __ldg(x4 + sub); Yn = __ldg(y4 + sub); Zn = __ldg(z4 + sub);
            if (pw) Wn = __ldg(w4 + sub);
        }
        for (int c = sub; c < nc; c += 8) {
            float4 X = Xn, Y = Yn, Z = Zn;
            const float4 W = Wn;
            if (c + 8 < nc) {
                Xn = __ldg(x4 + c + 8); Yn = __ldg(y4 + c + 8); Zn = __ldg(z4 + c + 8);
                if (pw) Wn = __ldg(w4 + c + 8);
            }
            if (R) {
                rotate(R, X.x, Y.x, Z.x); rotate(R, X.y, Y.y, Z.y);
                rotate(R, X.z, Y.z, Z.z); rotate(R, X.w, Y.w, Z.w);
            }
            acc_point<ORDER, PAR>(A, X.x, Y.x, Z.x, use_w ? W.x : 1.f, W.x);
            acc_point<ORDER, PAR>(A, X.y, Y.y, Z.y, use_w ? W.y : 1.f, W.y);
            acc_point<ORDER, PAR>(A, X.z, Y.z, Z.z, use_w ? W.z : 1.f, W.z);
            acc_point<ORDER, PAR>(A, X.w, Y.w, Z.w, use_w ? W.w : 1.f, W.w);
        }
    } else {
        for (int j = sub; j < k; j += 8) {
            float x = __ldg(px + j), y = __ldg(py + j), z = __ldg(pz + j);
            const float w = pw ? __ldg(pw + j) : 1.f;
            if (R) rotate(R, x, y, z);
            acc_point<ORDER, PAR>(A, x, y, z, use_w ? w : 1.f, w);
        }
    }
}

__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
__device__ __forceinline__ double group8_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// the group's totals of the moments this warp owns -> the patch's slot (lane j of the group stores every 8th of them)
template <int ORDER, int PAR>
__device__ __forceinline__ void reduce_store(Acc<ORDER>& A, double* slot, int sub) {
    using J = Jet<ORDER>;
    int n = 0;   // running count of owned entries (compile-time after unrolling)
#pragma unroll
    for (int a = 0; a <= J::D; ++a) {
        if (!owns_m<PAR>(a)) continue;
#pragma unroll
        for (int b = 0; b <= J::D - a; ++b) {
            const double t = group8_sum(A.m[J::mi(a, b)]);
            if ((n & 7) == sub) slot[J::mi(a, b)] = t;
            ++n;
        }
    }
#pragma unroll
    for (int a = 0; a <= ORDER; ++a) {
        if (!owns_z<PAR>(a)) continue;
#pragma unroll
        for (int b = 0; b <= ORDER - a; ++b) {
            const double t = group8_sum(A.zm[J::zi(a, b)]);
            if ((n & 7) == sub) slot[J::NM + J::zi(a, b)] = t;
            ++n;
        }
    }
    if (PAR == 1) {
        A.sax = group8_sum(A.sax);
        A.say = group8_sum(A.say);
        if (sub == 0) {
            slot[J::NM + J::NZ] = A.sax;
            slot[J::NM + J::NZ + 1] = A.say;
        }
    }
}

// In-register unblocked Cholesky (lower, packed).  Failure criterion as LAPACK spotrf: pivot <= 0 or NaN.
#define LIDX(i, j) ((i) * ((i) + 1) / 2 + (j))
template <int M>
__device__ __forceinline__ bool chol_factor(float (&a)[M * (M + 1) / 2]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        float d = a[LIDX(j, j)];
#pragma unroll
        for (int t = 0; t < j; ++t) d = fmaf(-a[LIDX(j, t)], a[LIDX(j, t)], d);
        if (!(d > 0.f)) ok = false;
        const float l = sqrtf(d);
        const float inv = 1.f / l;
        a[LIDX(j, j)] = l;
#pragma unroll
        for (int i = j + 1; i < M; ++i) {
            float s = a[LIDX(i, j)];
#pragma unroll
            for (int t = 0; t < j; ++t) s = fmaf(-a[LIDX(i, t)], a[LIDX(j, t)], s);
            a[LIDX(i, j)] = s * inv;
        }
    }
    return ok;
}
// Solve L L^T x = b in place (b -> x) with the factor from chol_factor.
template <int M>
__device__ __forceinline__ void chol_substitute(const float (&a)[M * (M + 1) / 2], float (&b)[M]) {
#pragma unroll
    for (int i = 0; i < M; ++i) {
        float s = b[i];
#pragma unroll
        for (int t = 0; t < i; ++t) s = fmaf(-a[LIDX(i, t)], b[t], s);
        b[i] = s / a[LIDX(i, i)];
    }
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
        float s = b[i];
#pragma unroll
        for (int t = i + 1; t < M; ++t) s = fmaf(-a[LIDX(t, i)], b[t], s);
        b[i] = s / a[LIDX(i, i)];
    }
}
#undef LIDX

// dirs (may be NULL): the 2 x 3 matrix compute_principal_curvatures returns next to the curvatures
// (tutorial_utils.py:223-224): the eigenvector matrix V of torch.symeig (eigenvectors in COLUMNS, column j for
// curvature j) with a zero third column appended, row-major.  LAPACK leaves the sign of each eigenvector open; here
// its largest-magnitude component is positive.
__device__ __forceinline__ void curvature_from_beta(float b0f, float b1f, float b2f, float b3f, float b4f,
                                                    float& k1, float& k2, float* dirs = nullptr) {
    // tutorial_utils.py:207-223.  M = -I^{-1} II is not symmetric; torch.symeig(upper=True) reads
    // (M00, M01, M11) only.  Evaluated in fp64 and rounded once.
    const double b0 = b0f, b1 = b1f, b2 = b2f, b3 = b3f, b4 = b4f;
    const double E = 1.0 + b0 * b0, G = 1.0 + b1 * b1, F = b1 * b0;
    const double nrm = sqrt(b0 * b0 + b1 * b1 + 1.0);
    const double e = 2.0 * b2 / nrm, f = b4 / nrm, g = 2.0 * b3 / nrm;
    const double det = E * G - F * F;
    const double m00 = -(G * e - F * f) / det;
    const double m01 = -(G * f - F * g) / det;
    const double m11 = -(E * g - F * f) / det;
    const double mid = 0.5 * (m00 + m11);
    const double hd = 0.5 * (m00 - m11);
    const double r = sqrt(hd * hd + m01 * m01);
    k1 = static_cast<float>(mid - r);
    k2 = static_cast<float>(mid + r);
    if (dirs) {
        // eigenvector of the larger eigenvalue mid + r of [[m00, m01], [m01, m11]]: (r + hd, m01) or (m01, r - hd),
        // whichever is longer (they are parallel); the other eigenvector is orthogonal to it
        double vx, vy;
        if (hd >= 0.0) { vx = r + hd; vy = m01; } else { vx = m01; vy = r - hd; }
        const double len = sqrt(vx * vx + vy * vy);
        if (len > 0.0) { vx /= len; vy /= len; } else { vx = 0.0; vy = 1.0; }   // umbilic: any basis
        double ux = vy, uy = -vx;                                                 // eigenvector of mid - r
        if (fabs(vx) >= fabs(vy) ? vx < 0.0 : vy < 0.0) { vx = -vx; vy = -vy; }
        if (fabs(ux) >= fabs(uy) ? ux < 0.0 : uy < 0.0) { ux = -ux; uy = -uy; }
        dirs[0] = (float)ux; dirs[1] = (float)vx; dirs[2] = 0.f;
        dirs[3] = (float)uy; dirs[4] = (float)vy; dirs[5] = 0.f;
    }
}

// row vector times trans^T, then times U^T (test_n_est.py:154-161, test_c_est.py:162-168); either may be NULL
__device__ __forceinline__ void rotate_back(const float* R, const float* U, float& vx, float& vy, float& vz) {
    if (R) {
        const float tx = vx * R[0] + vy * R[1] + vz * R[2];
        const float ty = vx * R[3] + vy * R[4] + vz * R[5];
        const float tz = vx * R[6] + vy * R[7] + vz * R[8];
        vx = tx; vy = ty; vz = tz;
    }
    if (U) {
        const float tx = vx * U[0] + vy * U[1] + vz * U[2];
        const float ty = vx * U[3] + vy * U[4] + vz * U[5];
        const float tz = vx * U[6] + vy * U[7] + vz * U[8];
        vx = tx; vy = ty; vz = tz;
    }
}

struct WlsArgs {
    const float* patch;
    const float* weights;
    const float* trans;
    const float* U;
    const double* rad;
    long long n;
    int k;
    float* beta;
    float* n_local;
    float* n_world;
    float* curv;
    double* curv_scaled;
    int* info;
    int keep_qstn;      // 1: do NOT undo the QSTN rotation on n_world / dirs_world (tutorial/compute_normals.py:67)
    float* dirs;        // [n,2,3] local frame (compute_principal_curvatures' second return value)
    float* dirs_world;  // [n,2,3] rows rotated back by trans^T and U^T (test_c_est.py:162-168)
    const int* out_row; // optional: patch i writes its outputs to row out_row[i] - out_base (queries visited in another order)
    long long out_base;
};

// Phase 2 for one patch (one lane): preconditioned normal equations from the patch's moment slot, Cholesky + two steps
// of fp64-residual refinement, D_inv, normal, curvatures, principal directions, back-rotation; outputs to row po.
template <int ORDER>
__device__ __forceinline__ void solve_and_emit(const WlsArgs& p, const double* slot, long long pi, long long po) {
    using J = Jet<ORDER>;
    constexpr int M = J::M;
    float h = 1.f;
    if (ORDER > 1) {
        // DeepFit.py:43-46 (fp32, like the reference)
        const float invk = 1.f / (float)p.k;
        h = ((float)slot[J::NM + J::NZ] * invk + (float)slot[J::NM + J::NZ + 1] * invk) * 0.5f;
        if (fabsf(h) < 1e-4f) h = 0.1f;
    }
    const double ih = 1.0 / (double)h;
    double ihp[J::D + 1];
    ihp[0] = 1.0;
#pragma unroll
    for (int a = 1; a <= J::D; ++a) ihp[a] = ihp[a - 1] * ih;

    double bsol[M];  // solution of the preconditioned system, refined in fp64
    int info = 0;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        float a[M * (M + 1) / 2];
        float rhs[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const int ax = colx<ORDER>(i) + colx<ORDER>(j), by = coly<ORDER>(i) + coly<ORDER>(j);
                double v = slot[J::mi(ax, by)] * ihp[ax + by];
                if (attempt == 1 && i == j) v *= 1.01;
                a[i * (i + 1) / 2 + j] = (float)v;
            }
            rhs[i] = (float)(slot[J::NM + J::zi(colx<ORDER>(i), coly<ORDER>(i))] * ihp[colx<ORDER>(i) + coly<ORDER>(i)]);
        }
        const bool ok = chol_factor<M>(a);
        if (!ok) {
            info = attempt + 1;
            continue;
        }
        chol_substitute<M>(a, rhs);
#pragma unroll
        for (int i = 0; i < M; ++i) bsol[i] = (double)rhs[i];
        // iterative refinement: residual of the (possibly ridged) fp64 system, correction through the
        // fp32 factor; contracts by ~cond * 2^-24 per step
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double r = slot[J::NM + J::zi(colx<ORDER>(i), coly<ORDER>(i))] * ihp[colx<ORDER>(i) + coly<ORDER>(i)];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    const int ax = colx<ORDER>(i) + colx<ORDER>(j), by = coly<ORDER>(i) + coly<ORDER>(j);
                    double v = slot[J::mi(ax, by)] * ihp[ax + by];
                    if (attempt == 1 && i == j) v *= 1.01;
                    r = fma(-v, bsol[j], r);
                }
                rhs[i] = (float)r;
            }
            chol_substitute<M>(a, rhs);
#pragma unroll
            for (int i = 0; i < M; ++i) bsol[i] += (double)rhs[i];
        }
        break;
    }
    float beta[M];
    // D_inv, DeepFit.py:85-86
#pragma unroll
    for (int i = 0; i < M; ++i)
        beta[i] = (info == 2) ? 0.f : (float)(bsol[i] * ihp[colx<ORDER>(i) + coly<ORDER>(i)]);

    if (p.beta) {
#pragma unroll
        for (int i = 0; i < M; ++i) p.beta[po * M + i] = beta[i];
    }
    // normal, DeepFit.py:88
    const float inv_n = 1.f / fmaxf(sqrtf(fmaf(beta[0], beta[0], fmaf(beta[1], beta[1], 1.f))), 1e-12f);
    const float nx = -beta[0] * inv_n, ny = -beta[1] * inv_n, nz = inv_n;
    if (p.n_local) {
        p.n_local[po * 3 + 0] = nx; p.n_local[po * 3 + 1] = ny; p.n_local[po * 3 + 2] = nz;
    }
    if (p.n_world) {
        float vx = nx, vy = ny, vz = nz;
        // n * trans^T (test_n_est.py:154-157), then n * U^T (compute_normals.py:67)
        rotate_back((p.trans && !p.keep_qstn) ? p.trans + pi * 9 : nullptr, p.U ? p.U + pi * 9 : nullptr, vx, vy, vz);
        p.n_world[po * 3 + 0] = vx; p.n_world[po * 3 + 1] = vy; p.n_world[po * 3 + 2] = vz;
    }
    if (ORDER > 1 && (p.curv || p.curv_scaled || p.dirs || p.dirs_world)) {
        float k1, k2, dirs[6];
        const bool want_dirs = p.dirs || p.dirs_world;
        curvature_from_beta(beta[0], beta[1], beta[2], beta[3], beta[4], k1, k2, want_dirs ? dirs : nullptr);
        if (p.curv) { p.curv[po * 2] = k1; p.curv[po * 2 + 1] = k2; }
        if (p.curv_scaled) {
            const double r = p.rad ? p.rad[pi] : 1.0;  // fp32 / fp64 -> fp64, compute_normals.py:79
            p.curv_scaled[po * 2] = (double)k1 / r;
            p.curv_scaled[po * 2 + 1] = (double)k2 / r;
        }
        if (p.dirs) {
#pragma unroll
            for (int i = 0; i < 6; ++i) p.dirs[po * 6 + i] = dirs[i];
        }
        if (p.dirs_world) {
#pragma unroll
            for (int rrow = 0; rrow < 2; ++rrow) {
                float vx = dirs[rrow * 3], vy = dirs[rrow * 3 + 1], vz = 0.f;
                rotate_back((p.trans && !p.keep_qstn) ? p.trans + pi * 9 : nullptr, p.U ? p.U + pi * 9 : nullptr, vx, vy, vz);
                p.dirs_world[po * 6 + rrow * 3] = vx; p.dirs_world[po * 6 + rrow * 3 + 1] = vy; p.dirs_world[po * 6 + rrow * 3 + 2] = vz;
            }
        }
    }
    if (p.info) p.info[po] = info;
}

// PPC = patches per CTA iteration (phase 1: both warps walk all of them, each for its half of the moments; phase 2: each
// warp solves PPC / 2 systems, one per lane): 64 when there is enough work to fill the GPU that way, 16 for small batches.
#ifndef DFB_WLS_MINB
#define DFB_WLS_MINB 8   // 128 registers: phase 1 needs exactly that at order 3; the per-lane solve of phase 2 spills a little
#endif
template <int ORDER, int PPC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, DFB_WLS_MINB) wls_kernel(WlsArgs p) {
    using J = Jet<ORDER>;
    constexpr int M = J::M;
    constexpr int SS = slot_stride<ORDER>();
    __shared__ double s_mom[PPC][SS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, sub = lane & 7;
    const long long n_super = (p.n + PPC - 1) / PPC;  // groups of PPC patches, one per CTA iteration

    for (long long sp = blockIdx.x; sp < n_super; sp += gridDim.x) {
        const long long base = sp * PPC;
        // ---------------- phase 1 ----------------
        for (int round = 0; round < PPC / 4; ++round) {
            long long pi = base + round * 4 + grp;
            const bool live = pi < p.n;
            if (!live) pi = p.n - 1;
            const float* px = p.patch + pi * 3LL * p.k;
            const float* pw = p.weights ? p.weights + pi * (long long)p.k : nullptr;
            float R[9];
            const bool has_R = p.trans != nullptr;
            if (has_R) {
#pragma unroll
                for (int i = 0; i < 9; ++i) R[i] = __ldg(p.trans + pi * 9 + i);
            }
            Acc<ORDER> A;
            // zero-weight guard, DeepFit.py:37-39: weights -> 1 unless more than 18 exceed 1e-3 (both warps count)
            if (warp == 0) {
                acc_patch<ORDER, 0>(A, px, pw, has_R ? R : nullptr, p.k, sub, true);
                bool redo = false;
                if (pw) redo = group8_sum(A.nvalid) <= 18.f;
                if (__any_sync(0xffffffffu, redo)) acc_patch<ORDER, 0>(A, px, pw, has_R ? R : nullptr, p.k, sub, !redo);
                reduce_store<ORDER, 0>(A, s_mom[round * 4 + grp], sub);
            } else {
                acc_patch<ORDER, 1>(A, px, pw, has_R ? R : nullptr, p.k, sub, true);
                bool redo = false;
                if (pw) redo = group8_sum(A.nvalid) <= 18.f;
                if (__any_sync(0xffffffffu, redo)) acc_patch<ORDER, 1>(A, px, pw, has_R ? R : nullptr, p.k, sub, !redo);
                reduce_store<ORDER, 1>(A, s_mom[round * 4 + grp], sub);
            }
        }
        __syncthreads();
        // ---------------- phase 2 ----------------
        const int sl = warp * (PPC / 2) + lane;   // this lane's slot of the batch
        const long long pi = base + sl;
        if (lane < PPC / 2 && pi < p.n)
            solve_and_emit<ORDER>(p, s_mom[sl], pi, p.out_row ? (long long)p.out_row[pi] - p.out_base : pi);
        __syncthreads();   // the slots are rewritten by the next iteration
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same kernel with its input STAGED THROUGH SHARED MEMORY by the TMA engine (cp.async.bulk + mbarrier): a ring of
// kStages stages, each one "round" = 4 patches x up to 256 points (x, y, z and weight rows, 1 KB bulk copies), filled
// two rounds ahead by one elected lane.  With register loads the warps sat on exposed DRAM latency (ncu: a quarter of
// the stall samples on the first use of the loaded chunk, FP64 pipe 43 % busy); here the bytes in flight per SM are
// set by the ring, not by registers and occupancy, both warps read each byte from shared memory (the second read cost
// an L1 / L2 request before), and patches longer than 256 points stream through the same stages chunk by chunk.
constexpr int kStages = 3;
constexpr int kChunkPts = 256;                         // points per patch and round
constexpr int kStageFloats = 4 * 4 * kChunkPts;        // 4 patches x (x, y, z, w) rows: 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded (2 s of wall clock): a protocol bug must end as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        if (mbar_try(bar, parity)) return;
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// one chunk (len points, a multiple of 4) of one patch from its staged rows; 8 lanes, sub = lane & 7
template <int ORDER, int PAR>
__device__ __forceinline__ void acc_chunk_smem(Acc<ORDER>& A, const float* sx, const float* sw, const float* R, int len, int sub,
                                               bool has_w) {
    const float4* x4 = reinterpret_cast<const float4*>(sx);
    const float4* y4 = reinterpret_cast<const float4*>(sx + kChunkPts);
    const float4* z4 = reinterpret_cast<const float4*>(sx + 2 * kChunkPts);
    const float4* w4 = reinterpret_cast<const float4*>(sw);
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    const int nc = len >> 2;
    float4 Xn = one4, Yn = one4, Zn = one4, Wn = one4;
    if (sub < nc) {
        Xn = x4[sub]; Yn = y4[sub]; Zn = z4[sub];
        if (has_w) Wn = w4[sub];
    }
    for (int c = sub; c < nc; c += 8) {
        float4 X = Xn, Y = Yn, Z = Zn;
        const float4 W = Wn;
        if (c + 8 < nc) {
            Xn = x4[c + 8]; Yn = y4[c + 8]; Zn = z4[c + 8];
            if (has_w) Wn = w4[c + 8];
        }
        if (R) {
            rotate(R, X.x, Y.x, Z.x); rotate(R, X.y, Y.y, Z.y);
            rotate(R, X.z, Y.z, Z.z); rotate(R, X.w, Y.w, Z.w);
        }
        acc_point<ORDER, PAR>(A, X.x, Y.x, Z.x, W.x, W.x);
        acc_point<ORDER, PAR>(A, X.y, Y.y, Z.y, W.y, W.y);
        acc_point<ORDER, PAR>(A, X.z, Y.z, Z.z, W.z, W.z);
        acc_point<ORDER, PAR>(A, X.w, Y.w, Z.w, W.w, W.w);
    }
}

template <int ORDER, int PAR>
__device__ __forceinline__ void staged_phase1(const WlsArgs& p, double (*s_mom)[slot_stride<ORDER>()], const float* ring,
                                              uint64_t* full, uint64_t* empty, long long base, int ppc, int nch, long long& g,
                                              int lane) {
    const int grp = lane >> 3, sub = lane & 7;
    const bool has_R = p.trans != nullptr, has_w = p.weights != nullptr;
    for (int round = 0; round < ppc / 4; ++round) {
        long long pi = base + round * 4 + grp;
        if (pi >= p.n) pi = p.n - 1;
        float R[9];
        if (has_R) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = __ldg(p.trans + pi * 9 + i);
        }
        Acc<ORDER> A;
#pragma unroll
        for (int i = 0; i < Jet<ORDER>::NM; ++i) A.m[i] = 0.0;
#pragma unroll
        for (int i = 0; i < Jet<ORDER>::NZ; ++i) A.zm[i] = 0.0;
        A.sax = A.say = A.nvalid = 0.f;
        for (int c = 0; c < nch; ++c, ++g) {
            const int st = (int)(g % kStages);
            mbar_wait(smem_u32(&full[st]), (uint32_t)(g / kStages) & 1u);
            const float* sp = ring + (size_t)st * kStageFloats + grp * (4 * kChunkPts);
            const int len = min(kChunkPts, p.k - c * kChunkPts);
            acc_chunk_smem<ORDER, PAR>(A, sp, sp + 3 * kChunkPts, has_R ? R : nullptr, len, sub, has_w);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty[st]));
        }
        // zero-weight guard, DeepFit.py:37-39 (weights -> 1 unless more than 18 exceed 1e-3): practically never taken
        // (sigmoid / tanh weights are >= 0.005), so the affected patches are simply accumulated again from global memory
        if (has_w) {
            const bool redo = group8_sum(A.nvalid) <= 18.f;
            if (__any_sync(0xffffffffu, redo))
                acc_patch<ORDER, PAR>(A, p.patch + pi * 3LL * p.k, p.weights + pi * (long long)p.k, has_R ? R : nullptr, p.k, sub, !redo);
        }
        reduce_store<ORDER, PAR>(A, s_mom[round * 4 + grp], sub);
    }
}

template <int ORDER, int PPC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, DFB_WLS_MINB) wls_kernel_staged(WlsArgs p) {
    constexpr int SS = slot_stride<ORDER>();
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    float* ring = reinterpret_cast<float*>(dyn_smem);
    __shared__ double s_mom[PPC][SS];
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_super = (p.n + PPC - 1) / PPC;
    const int nch = (p.k + kChunkPts - 1) / kChunkPts;           // chunks per patch
    const int rounds_per_it = (PPC / 4) * nch;
    const long long n_it = blockIdx.x < n_super ? (n_super - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_rounds = n_it * rounds_per_it;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(smem_u32(&full[i]), 1);
            mbar_init(smem_u32(&empty[i]), kWarpsPerBlock);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool has_w = p.weights != nullptr;
    // the producer: round r of this CTA -> (iteration, patch quad, chunk) -> 12 or 16 row copies of <= 1 KB
    auto issue = [&](long long r) {
        const int st = (int)(r % kStages);
        if (r >= kStages) mbar_wait(smem_u32(&empty[st]), (uint32_t)(r / kStages - 1) & 1u);
        const long long it = r / rounds_per_it;
        const int rr = (int)(r % rounds_per_it);
        const int quad = rr / nch, c = rr % nch;
        const long long base = ((long long)blockIdx.x + it * gridDim.x) * PPC + quad * 4;
        const int len = min(kChunkPts, p.k - c * kChunkPts);
        const uint32_t bytes = (uint32_t)len * 4u;
        const uint32_t bar = smem_u32(&full[st]);
        mbar_expect_tx(bar, bytes * 4u * (has_w ? 4u : 3u));
        const uint32_t dst0 = smem_u32(ring + (size_t)st * kStageFloats);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            long long pi = base + j;
            if (pi >= p.n) pi = p.n - 1;
            const float* src = p.patch + pi * 3LL * p.k + c * kChunkPts;
            const uint32_t dst = dst0 + (uint32_t)(j * 4 * kChunkPts * 4);
            bulk_g2s(dst, src, bytes, bar);
            bulk_g2s(dst + kChunkPts * 4, src + p.k, bytes, bar);
            bulk_g2s(dst + 2 * kChunkPts * 4, src + 2 * p.k, bytes, bar);
            if (has_w) bulk_g2s(dst + 3 * kChunkPts * 4, p.weights + pi * (long long)p.k + c * kChunkPts, bytes, bar);
        }
    };
    const bool producer = threadIdx.x == 0;
    if (producer)
        for (long long r = 0; r < kStages - 1 && r < n_rounds; ++r) issue(r);
    long long g = 0;        // rounds consumed by this warp
    long long issued = kStages - 1;
    for (long long it = 0; it < n_it; ++it) {
        const long long base = ((long long)blockIdx.x + it * gridDim.x) * PPC;
        // ---------------- phase 1 (the producer lane tops the ring up before each of its rounds) ----------------
        // The refill is interleaved at quad granularity: before quad q of this iteration the producer issues every round
        // up to (rounds consumed so far) + kStages - 1.
        for (int quad = 0; quad < PPC / 4; ++quad) {
            if (producer) {
                const long long upto = min(n_rounds, g + kStages);   // this warp has consumed every round below g
                for (; issued < upto; ++issued) issue(issued);
            }
            __syncwarp();
            if (warp == 0) staged_phase1<ORDER, 0>(p, s_mom + quad * 4, ring, full, empty, base + quad * 4, 4, nch, g, lane);
            else staged_phase1<ORDER, 1>(p, s_mom + quad * 4, ring, full, empty, base + quad * 4, 4, nch, g, lane);
        }
        __syncthreads();
        // ---------------- phase 2 ----------------
        const int sl = warp * (PPC / 2) + lane;
        const long long pi = base + sl;
        if (lane < PPC / 2 && pi < p.n)
            solve_and_emit<ORDER>(p, s_mom[sl], pi, p.out_row ? (long long)p.out_row[pi] - p.out_base : pi);
        __syncthreads();
    }
}

__global__ void curvature_kernel(const float* __restrict__ beta, long long n, int m, float* __restrict__ curv,
                                 float* __restrict__ dirs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = beta + i * m;
    float k1, k2, d[6];
    curvature_from_beta(b[0], b[1], b[2], b[3], b[4], k1, k2, dirs ? d : nullptr);
    if (curv) {
        curv[i * 2] = k1;
        curv[i * 2 + 1] = k2;
    }
    if (dirs) {
#pragma unroll
        for (int j = 0; j < 6; ++j) dirs[i * 6 + j] = d[j];
    }
}

// Analytic normals of the fitted jet at every neighbour (fit_Wjet(..., compute_neighbor_normals=True),
// reference models/DeepFit.py:90-115).  The reference evaluates the gradient with the PRECONDITIONED
// coordinates x/h, y/h but the un-preconditioned beta; that quirk is reproduced.  One warp per patch.
__global__ void __launch_bounds__(256) neighbor_normals_kernel(const float* __restrict__ patch, const float* __restrict__ trans,
                                                               const float* __restrict__ beta, long long n, int k, int order,
                                                               float* __restrict__ out) {
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= n) return;
    const float* px = patch + p * 3LL * k;
    float R[9];
    if (trans) {
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(trans + p * 9 + i);
    }
    float sax = 0.f, say = 0.f;
    for (int j = lane; j < k; j += 32) {
        float x = px[j], y = px[k + j], z = px[2 * k + j];
        if (trans) rotate(R, x, y, z);
        sax += fabsf(x);
        say += fabsf(y);
    }
    sax = warp_sum(sax);
    say = warp_sum(say);
    float h = 1.f;
    if (order > 1) {
        h = (sax / (float)k + say / (float)k) * 0.5f;
        if (fabsf(h) < 1e-4f) h = 0.1f;
    }
    const int M = (order + 1) * (order + 2) / 2;
    float b[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) b[i] = i < M ? __ldg(beta + p * M + i) : 0.f;
    for (int j = lane; j < k; j += 32) {
        float x = px[j], y = px[k + j], z = px[2 * k + j];
        if (trans) rotate(R, x, y, z);
        if (order > 1) { x = x / h; y = y / h; }
        float gx = b[0], gy = b[1];
        if (order >= 2) {
            gx += 2.f * b[2] * x + b[4] * y;
            gy += 2.f * b[3] * y + b[4] * x;
        }
        if (order >= 3) {
            gx += 3.f * b[5] * x * x + 2.f * b[7] * x * y + b[8] * y * y;
            gy += 3.f * b[6] * y * y + b[7] * x * x + 2.f * b[8] * x * y;
        }
        if (order >= 4) {
            gx += 4.f * b[9] * x * x * x + 3.f * b[11] * x * x * y + b[12] * y * y * y + 2.f * b[13] * y * y * x;
            gy += 4.f * b[10] * y * y * y + b[11] * x * x * x + 3.f * b[12] * x * y * y + 2.f * b[13] * y * x * x;
        }
        const float inv = 1.f / fmaxf(sqrtf(gx * gx + gy * gy + 1.f), 1e-12f);
        float* o = out + (p * k + j) * 3;
        o[0] = -gx * inv; o[1] = -gy * inv; o[2] = inv;
    }
}

template <int ORDER>
int launch_wls(const WlsArgs& a, cudaStream_t st) {
    // 64 patches per CTA iteration when that still gives every SM several CTAs' worth of iterations; 16 otherwise (a
    // 16 K-patch chunk of the DeepFit path): phase 2 then runs on a quarter of the lanes, but the GPU is filled
    const bool small = (a.n + 63) / 64 < (long long)sm_count() * 16;
    const int ppc = small ? 16 : 64;
    const long long n_super = (a.n + ppc - 1) / ppc;
    long long blocks = n_super;
    const long long cap = (long long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    // TMA-staged input: measured (profiles/r2_wls_staged_ab.log) 20 % faster at order 1, where the kernel waits on its
    // loads, and 10-50 % slower at orders 2-4, where the FP64 pipe is the limiter and the ring's 48 KB cost resident warps:
    // default for order 1 only (DFB_WLS_STAGED=0 / 1 forces either kernel for A/B runs).  Bulk copies need 16-byte
    // aligned rows, and the ring holds at most kStages chunks of a patch.
    static const int force_staged = [] { const char* e = getenv("DFB_WLS_STAGED"); return e ? (e[0] == '1' ? 1 : 0) : -1; }();
    const bool want_staged = force_staged >= 0 ? force_staged == 1 : ORDER == 1;
    const bool staged = want_staged && a.k % 4 == 0 && a.k <= kChunkPts * (kStages - 1) &&
                        (reinterpret_cast<uintptr_t>(a.patch) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.weights) & 15) == 0;
    if (staged) {
        constexpr int kRingBytes = kStages * kStageFloats * 4;
        if (small) {
            cudaFuncSetAttribute(wls_kernel_staged<ORDER, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
            wls_kernel_staged<ORDER, 16><<<(unsigned)blocks, kWarpsPerBlock * 32, kRingBytes, st>>>(a);
        } else {
            cudaFuncSetAttribute(wls_kernel_staged<ORDER, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes);
            wls_kernel_staged<ORDER, 64><<<(unsigned)blocks, kWarpsPerBlock * 32, kRingBytes, st>>>(a);
        }
    } else if (small) wls_kernel<ORDER, 16><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(a);
    else wls_kernel<ORDER, 64><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(a);
    note_launch();
    return check_cuda(cudaGetLastError(), "wls_kernel launch", __FILE__, __LINE__);
}

}  // namespace
}  // namespace dfb

namespace dfb {
int wls_fit_impl(const float* d_patch, const float* d_weights, const float* d_trans, const float* d_U, const double* d_rad,
                 int64_t n, int32_t k, int32_t order, float* d_beta, float* d_n_local, float* d_n_world, float* d_curv,
                 double* d_curv_scaled, float* d_dirs, float* d_dirs_world, int32_t* d_info, int keep_qstn, dfb_stream stream,
                 const int32_t* d_out_row, int64_t out_base) {
    DFB_REQUIRE(order >= 1 && order <= 4, DFB_E_ARG, "Polynomial order unsupported, please use 1..4 (got %d)", order);
    DFB_REQUIRE(n >= 0 && k >= 1, DFB_E_ARG, "dfb_wls_fit: bad sizes n=%lld k=%d", (long long)n, k);
    DFB_REQUIRE(d_patch != nullptr || n == 0, DFB_E_ARG, "dfb_wls_fit: d_patch is NULL");
    DFB_REQUIRE(order > 1 || (d_curv == nullptr && d_curv_scaled == nullptr && d_dirs == nullptr && d_dirs_world == nullptr),
                DFB_E_ARG, "Can't compute curvatures for jet with order less than 2");
    if (n == 0) return DFB_OK;
    int rc = dfb_device_check();
    if (rc) return rc;
    WlsArgs a{d_patch, d_weights, d_trans, d_U, d_rad, (long long)n, k,
              d_beta,  d_n_local, d_n_world, d_curv, d_curv_scaled, d_info, keep_qstn, d_dirs, d_dirs_world,
              d_out_row, (long long)out_base};
    switch (order) {
        case 1: return launch_wls<1>(a, as_stream(stream));
        case 2: return launch_wls<2>(a, as_stream(stream));
        case 3: return launch_wls<3>(a, as_stream(stream));
        default: return launch_wls<4>(a, as_stream(stream));
    }
}
}  // namespace dfb

extern "C" int dfb_wls_fit(const float* d_patch, const float* d_weights, const float* d_trans, const float* d_U,
                           const double* d_rad, int64_t n, int32_t k, int32_t order, float* d_beta,
                           float* d_n_local, float* d_n_world, float* d_curv, double* d_curv_scaled,
                           float* d_dirs, float* d_dirs_world, int32_t* d_info, dfb_stream stream) {
    return dfb::wls_fit_impl(d_patch, d_weights, d_trans, d_U, d_rad, n, k, order, d_beta, d_n_local, d_n_world, d_curv,
                             d_curv_scaled, d_dirs, d_dirs_world, d_info, 0, stream, nullptr, 0);
}

extern "C" int dfb_principal_curvatures(const float* d_beta, int64_t n, int32_t m, float* d_curv, float* d_dirs,
                                        dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(m >= 5, DFB_E_ARG, "Can't compute curvatures for jet with order less than 2");
    DFB_REQUIRE(n >= 0 && (n == 0 || (d_beta && (d_curv || d_dirs))), DFB_E_ARG, "dfb_principal_curvatures: bad arguments");
    if (n == 0) return DFB_OK;
    int rc = dfb_device_check();
    if (rc) return rc;
    curvature_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(d_beta, (long long)n, m, d_curv, d_dirs);
    note_launch();
    return check_cuda(cudaGetLastError(), "curvature_kernel launch", __FILE__, __LINE__);
}

extern "C" int dfb_neighbor_normals(const float* d_patch, const float* d_trans, const float* d_beta, int64_t n, int32_t k,
                                    int32_t order, float* d_out, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(order >= 1 && order <= 4, DFB_E_ARG, "Polynomial order unsupported, please use 1..4 (got %d)", order);
    DFB_REQUIRE(n >= 0 && k >= 1 && (n == 0 || (d_patch && d_beta && d_out)), DFB_E_ARG, "dfb_neighbor_normals: bad arguments");
    if (n == 0) return DFB_OK;
    int rc = dfb_device_check();
    if (rc) return rc;
    neighbor_normals_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(d_patch, d_trans, d_beta, (long long)n, k, order, d_out);
    note_launch();
    return check_cuda(cudaGetLastError(), "neighbor_normals_kernel launch", __FILE__, __LINE__);
}
