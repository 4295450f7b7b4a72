// Stage 1: exact k-nearest-neighbour search on a Morton-ordered multi-level uniform grid.
//
// Replaces scipy.spatial.cKDTree(points, 10) + query(points[i], k) (reference
// tutorial/tutorial_utils.py:16,21; dataset.py:32,293).  Results are bit-identical to cKDTree after
// distance-then-index tie-breaking: the ranking key is (fp64 d^2, index) with
// d^2 = ((double)px-(double)qx)^2 + (..y)^2 + (..z)^2 accumulated in that order without FMA
// contraction, exactly what cKDTree evaluates on float32 inputs promoted to double.
//
// Structure
//   build : bounding cube -> 2^b cells per axis; points counting-sorted by the Morton code of their
//           finest cell, so a cell of ANY coarser level l is one contiguous range
//           [tab[m << 3l], tab[(m+1) << 3l]) of the sorted array -- one table serves all levels.
//   query : one warp per query.  Climb levels until the 3x3x3 block of cells around the query holds
//           >= 3k points, stream the block's points (coalesced float4 reads), keep the k best in a
//           shared-memory list + insertion buffer (bitonic sort when the buffer fills, cells farther
//           than the current k-th distance are skipped), and accept when the k-th distance is
//           strictly inside the block's guaranteed margin; otherwise retry one level up with the
//           k-th key as the starting threshold.  The top level covers the whole cloud, so the
//           search always terminates with the exact answer.
#include "common.cuh"

#include <math.h>

namespace dfb {
namespace {

constexpr int kMaxLevelBits = 10;
constexpr float kCellSlack = 2e-3f;  // bound (in cells) on the float error of cell coordinates

__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

struct GridParams {
    float minx, miny, minz;
    float inv_cell;  // cells per unit length at the finest level
    float cell;      // finest cell edge
    int bits;        // finest level has 2^bits cells per axis
    long long n;
};

// Continuous cell coordinate of a point at the finest level.  Build and query share this function,
// so cell membership is consistent and monotone in the coordinate.
__device__ __forceinline__ float cell_coord(float p, float mn, float inv_cell) { return (p - mn) * inv_cell; }
__device__ __forceinline__ int cell_index(float u, int g) {
    int i = (int)floorf(u);
    return i < 0 ? 0 : (i >= g ? g - 1 : i);
}

// ---------------------------------------------------------------------------------------- build

__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

__global__ void bbox_kernel(const float* __restrict__ pts, long long n, unsigned* __restrict__ out6) {
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned o = f2ord(pts[i * 3 + a]);
            mn[a] = min(mn[a], o);
            mx[a] = max(mx[a], o);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(out6 + a, mn[a]);
            atomicMax(out6 + 3 + a, mx[a]);
        }
    }
}

// Bounding cube -> grid parameters, entirely on the device so (re)builds never synchronise the host.
__global__ void params_kernel(unsigned* __restrict__ bb6, GridParams* __restrict__ gp, int bits, long long n) {
    float mn[3], ext = 0.f;
    for (int a = 0; a < 3; ++a) {
        mn[a] = ord2f(bb6[a]);
        const float mx = ord2f(bb6[3 + a]);
        ext = fmaxf(ext, mx - mn[a]);
    }
    if (!(ext > 0.f) || !isfinite(ext)) ext = 1.f;
    ext *= 1.0001f;
    gp->minx = mn[0]; gp->miny = mn[1]; gp->minz = mn[2];
    gp->bits = bits;
    gp->inv_cell = (float)(1 << bits) / ext;
    gp->cell = ext / (float)(1 << bits);
    gp->n = n;
    for (int a = 0; a < 3; ++a) { bb6[a] = 0xffffffffu; bb6[3 + a] = 0u; }  // re-arm for the next build
}

__global__ void count_kernel(const float* __restrict__ pts, const GridParams* __restrict__ gp, uint32_t* __restrict__ code,
                             uint32_t* __restrict__ tab) {
    const GridParams g = *gp;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) return;
    const int G = 1 << g.bits;
    const uint32_t m = morton3(cell_index(cell_coord(pts[i * 3 + 0], g.minx, g.inv_cell), G),
                               cell_index(cell_coord(pts[i * 3 + 1], g.miny, g.inv_cell), G),
                               cell_index(cell_coord(pts[i * 3 + 2], g.minz, g.inv_cell), G));
    code[i] = m;
    atomicAdd(tab + 1 + m, 1u);
}

// Exclusive scan of tab[1..ncell] in place (tab[0] stays 0 and tab[c+1] becomes start of cell c,
// then the scatter turns it into the end of cell c == start of cell c+1).
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void scan_tile_sums(const uint32_t* __restrict__ v, long long n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t s[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile;
    uint32_t t = 0;
    for (int j = 0; j < kScanItems; ++j) {
        const long long i = base + (long long)j * kScanThreads + threadIdx.x;
        if (i < n) t += v[i];
    }
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t a = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) a += s[w];
        sums[blockIdx.x] = a;
    }
}

__global__ void scan_sums_serial(uint32_t* __restrict__ sums, long long m) {
    // single block; m <= 2^30 / 4096 = 262144
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_w[32];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (long long base = 0; base < m; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const uint32_t v = i < m ? sums[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? s_w[threadIdx.x] : 0u;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_w[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_off = (threadIdx.x >> 5) ? s_w[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t incl = x + warp_off + s_carry;
        if (i < m) sums[i] = incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = incl;
        __syncthreads();
    }
}

__global__ void scan_apply(uint32_t* __restrict__ v, long long n, const uint32_t* __restrict__ sums) {
    // exclusive scan inside each tile: thread-strided layout -> do it via shared memory in item order
    __shared__ uint32_t s[kScanTile];
    __shared__ uint32_t s_w[kScanThreads / 32];
    const long long base = (long long)blockIdx.x * kScanTile;
    for (int j = 0; j < kScanItems; ++j) {
        const int li = j * kScanThreads + threadIdx.x;
        s[li] = (base + li < n) ? v[base + li] : 0u;
    }
    __syncthreads();
    // each thread owns kScanItems consecutive items
    uint32_t loc[kScanItems];
    uint32_t t = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
        loc[j] = t;
        t += s[threadIdx.x * kScanItems + j];
    }
    uint32_t x = t;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t off = sums[blockIdx.x] + (x - t);
    for (int w = 0; w < (threadIdx.x >> 5); ++w) off += s_w[w];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) s[threadIdx.x * kScanItems + j] = loc[j] + off;
    __syncthreads();
    for (int j = 0; j < kScanItems; ++j) {
        const int li = j * kScanThreads + threadIdx.x;
        if (base + li < n) v[base + li] = s[li];
    }
}

__global__ void scatter_kernel(const float* __restrict__ pts, long long n, const uint32_t* __restrict__ code,
                               uint32_t* __restrict__ tab, float4* __restrict__ sorted) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pos = atomicAdd(tab + 1 + code[i], 1u);
    sorted[pos] = make_float4(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], __int_as_float((int)i));
}

// ---------------------------------------------------------------------------------------- query

struct Key {
    double d2;
    int idx;
};
__device__ __forceinline__ bool key_less(double a, int ai, double b, int bi) { return a < b || (a == b && ai < bi); }

// Ascending bitonic sort of cap (power of two) entries held in shared memory by one warp.
__device__ __forceinline__ void warp_bitonic_sort(double* sd, int* si, int cap, int lane) {
    for (int size = 2; size <= cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (cap >> 1); t += 32) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const double a = sd[lo], b = sd[hi];
                const int ai = si[lo], bi = si[hi];
                const bool sw = up ? key_less(b, bi, a, ai) : key_less(a, ai, b, bi);
                if (sw) {
                    sd[lo] = b; sd[hi] = a;
                    si[lo] = bi; si[hi] = ai;
                }
            }
            __syncwarp();
        }
    }
}

// Tighten the list without sorting it: find, by bisection, a bound T with at least k (and at most a few more) of the
// count entries <= T, then compact those to the front.  T is an UPPER bound of the k-th smallest key seen so far, which
// is all the pruning needs (every true neighbour has d2 <= T; ties are kept, the final sort orders them).  The bisection
// runs on the BIT PATTERNS of the (non-negative) distances: they are monotone in the value and roughly logarithmic, so a
// list that mixes a dense cluster with far points converges as fast as a uniform one, and as an integer binary search it
// ends at the exact k-th value at the latest.  ~1 k instructions where a 512-entry bitonic sort costs ~7 k.
// Returns the new count (>= k; larger than k + 16 only for exact ties).  EXACT: run the search to its end, T is then
// the k-th smallest distance itself and the new count is k unless distances tie across the cut.
template <bool EXACT>
__device__ __forceinline__ int warp_select_compact(double* sd, int* si, int count, int k, int lane, double& T) {
    long long lo = 0x7ff0000000000000LL, hi = 0;
    for (int t = lane; t < count; t += 32) {
        const long long v = __double_as_longlong(sd[t]);
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    // invariant: #(<= hi) >= k and #(< lo) < k
    for (int iter = 0; iter < 64 && lo < hi; ++iter) {
        const long long mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int t = lane; t < count; t += 32) c += (__double_as_longlong(sd[t]) <= mid) ? 1 : 0;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (c >= k) {
            hi = mid;
            if (!EXACT && c <= k + 16) break;
        } else {
            lo = mid + 1;
        }
    }
    T = __longlong_as_double(hi);
    int n = 0;
    for (int base = 0; base < count; base += 32) {
        const int t = base + lane;
        double v = 0.0;
        int id = 0;
        bool keep = false;
        if (t < count) {
            v = sd[t];
            id = si[t];
            keep = __double_as_longlong(v) <= hi;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) {
            const int pos = n + __popc(bal & ((1u << lane) - 1u));
            sd[pos] = v;
            si[pos] = id;
        }
        n += __popc(bal);
        __syncwarp();
    }
    return n;
}

struct KnnArgs {
    const float4* sorted;
    const uint32_t* tab;
    const float* pts;  // original order, f32[n,3]
    const GridParams* gp;
    const int* query;
    long long q_begin, q_count;
    int k, kcap;  // kcap = capacity of the shared list (power of two >= 2k, >= 64)
    int* out_idx;
    double* out_dist;
    double* out_rad;
};

constexpr int kKnnWarps = 4;

// Visiting order of the 27 cells of a block: the query's own cell first, then the 6 face, 12 edge and
// 8 corner neighbours.  Near cells first tightens the k-th-distance threshold early, so most later
// candidates are rejected before insertion and far cells are skipped whole.  Index = x + 3y + 9z.
__device__ __constant__ unsigned char c_cell_order[27] = {13, 12, 14, 10, 16, 4, 22, 9, 11, 15, 17, 3, 5, 21, 23, 1, 7, 19, 25,
                                                          0, 2, 6, 8, 18, 20, 24, 26};

__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(KnnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* sd = reinterpret_cast<double*>(smem_raw) + (size_t)warp * a.kcap;
    int* si = reinterpret_cast<int*>(smem_raw + (size_t)kKnnWarps * a.kcap * sizeof(double)) + (size_t)warp * a.kcap;
    const GridParams gpar = *a.gp;
    const int G = 1 << gpar.bits;
    const int k = a.k, cap = a.kcap;
    int k2 = 32;               // size of the sorting network for exactly k entries
    while (k2 < k) k2 <<= 1;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int level_hint = 0;   // the level that answered the warp's previous query: the search starts there, first DOWN while
                          // the finer block still holds enough points (a query that jumps from a sparse to a dense
                          // region must not scan a coarse block), then up as usual.  Any level passing the margin test
                          // is exact.

    for (long long qw = (long long)blockIdx.x * kKnnWarps + warp; qw < a.q_count; qw += (long long)gridDim.x * kKnnWarps) {
        const long long qi = a.query ? (long long)a.query[qw] : a.q_begin + qw;
        const float qxf = a.pts[qi * 3], qyf = a.pts[qi * 3 + 1], qzf = a.pts[qi * 3 + 2];
        const double qx = qxf, qy = qyf, qz = qzf;
        const float ux = cell_coord(qxf, gpar.minx, gpar.inv_cell), uy = cell_coord(qyf, gpar.miny, gpar.inv_cell),
                    uz = cell_coord(qzf, gpar.minz, gpar.inv_cell);
        const int ix = cell_index(ux, G), iy = cell_index(uy, G), iz = cell_index(uz, G);

        // threshold: candidates must satisfy key <= T when thr_incl (carried over from a failed
        // level) or key < T otherwise (T = current k-th best of the list)
        double Td = INF;
        int Ti = 0x7fffffff;
        bool thr_incl = true;

        const long long need = min((long long)3 * k, gpar.n);
        int l0 = level_hint;
        while (l0 > 0) {   // population of the 27-cell block one level finer
            const int lf = l0 - 1;
            const int Gf = G >> lf;
            const int fx = ix >> lf, fy = iy >> lf, fz = iz >> lf;
            unsigned cnt = 0;
            if (lane < 27) {
                const int ex = fx + (lane % 3) - 1, ey = fy + ((lane / 3) % 3) - 1, ez = fz + (lane / 9) - 1;
                if (ex >= 0 && ey >= 0 && ez >= 0 && ex < Gf && ey < Gf && ez < Gf) {
                    const uint32_t m = morton3(ex, ey, ez);
                    const int sh = 3 * lf;
                    cnt = a.tab[((size_t)m + 1) << sh] - a.tab[(size_t)m << sh];
                }
            }
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if ((long long)cnt < need) break;
            l0 = lf;
        }
        for (int l = l0; l <= gpar.bits; ++l) {
            const int Gl = G >> l;
            const int cx = ix >> l, cy = iy >> l, cz = iz >> l;
            // 27 neighbour cells, one per lane
            uint32_t lo = 0, hi = 0;
            float cmin2 = 0.f;  // conservative squared distance from the query to the cell (world units)
            if (lane < 27) {
                const int ex = cx + (lane % 3) - 1, ey = cy + ((lane / 3) % 3) - 1, ez = cz + (lane / 9) - 1;
                if (ex >= 0 && ey >= 0 && ez >= 0 && ex < Gl && ey < Gl && ez < Gl) {
                    const uint32_t m = morton3(ex, ey, ez);
                    const int sh = 3 * l;
                    if (sh >= 30) {  // single top-level cell
                        lo = 0;
                        hi = (uint32_t)gpar.n;
                    } else {
                        lo = a.tab[(size_t)m << sh];
                        hi = a.tab[((size_t)m + 1) << sh];
                    }
                    const float w = (float)(1 << l);
                    float dd[3];
                    const float uu[3] = {ux, uy, uz};
                    const int ee[3] = {ex, ey, ez};
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const float c0 = (float)ee[t] * w, c1 = c0 + w;
                        float d = 0.f;
                        if (uu[t] < c0) d = c0 - uu[t] - kCellSlack;
                        else if (uu[t] > c1) d = uu[t] - c1 - kCellSlack;
                        dd[t] = fmaxf(d, 0.f);
                    }
                    const float dc = gpar.cell * (1.f - 1e-5f);
                    cmin2 = (dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]) * dc * dc * (1.f - 1e-5f);
                }
            }
            unsigned total = hi - lo;
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            const bool top = (l == gpar.bits);
            if (!top && (long long)total < need) continue;

            // guaranteed-complete radius of the block (world units), strict
            double margin2 = INF;
            if (!top) {
                const float w = (float)(1 << l);
                float mrg = 3.0e38f;
                const float uu[3] = {ux, uy, uz};
                const int cc[3] = {cx, cy, cz};
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    if (cc[t] - 1 > 0) mrg = fminf(mrg, uu[t] - (float)(cc[t] - 1) * w - kCellSlack);
                    if (cc[t] + 2 < Gl) mrg = fminf(mrg, (float)(cc[t] + 2) * w - uu[t] - kCellSlack);
                }
                if (mrg < 3.0e38f) {
                    const double mm = (double)fmaxf(mrg, 0.f) * (double)gpar.cell * (1.0 - 1e-6);
                    margin2 = mm * mm;
                }
            }

            // (re)start the list
            int count = 0;
            bool full_once = false;  // list has been cut to k at least once => threshold is strict on list[k-1]
            for (int ci = 0; ci < 27; ++ci) {
                const int c = c_cell_order[ci];
                const uint32_t clo = __shfl_sync(0xffffffffu, lo, c), chi = __shfl_sync(0xffffffffu, hi, c);
                const float cm2 = __shfl_sync(0xffffffffu, cmin2, c);
                if (chi <= clo) continue;
                if ((double)cm2 > Td) continue;
                for (uint32_t j0 = clo; j0 < chi; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    bool take = false;
                    double d2 = 0.0;
                    int pidx = 0;
                    if (j < chi) {
                        const float4 p = __ldg(a.sorted + j);
                        const double dx = (double)p.x - qx, dy = (double)p.y - qy, dz = (double)p.z - qz;
                        d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        pidx = __float_as_int(p.w);
                        take = thr_incl ? !key_less(Td, Ti, d2, pidx) : key_less(d2, pidx, Td, Ti);
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, take);
                    if (bal == 0u) continue;
                    if (count + 32 > cap) {
                        // the list is full: tighten the bound by selection; only if that cannot make room (a pile of
                        // exact ties) pad, sort and keep exactly the k best
                        double T;
                        count = warp_select_compact<false>(sd, si, count, k, lane, T);
                        Td = T;
                        Ti = 0x7fffffff;
                        thr_incl = true;
                        full_once = true;
                        if (count + 32 > cap) {
                            for (int t = count + lane; t < cap; t += 32) { sd[t] = INF; si[t] = 0x7fffffff; }
                            __syncwarp();
                            warp_bitonic_sort(sd, si, cap, lane);
                            count = k;
                            Td = sd[k - 1];
                            Ti = si[k - 1];
                            thr_incl = false;
                            __syncwarp();
                        }
                        // re-evaluate this batch against the tighter threshold
                        take = (j < chi) && (thr_incl ? !key_less(Td, Ti, d2, pidx) : key_less(d2, pidx, Td, Ti));
                        const unsigned bal2 = __ballot_sync(0xffffffffu, take);
                        if (bal2 == 0u) continue;
                        const int pos = count + __popc(bal2 & ((1u << lane) - 1u));
                        if (take) { sd[pos] = d2; si[pos] = pidx; }
                        count += __popc(bal2);
                        __syncwarp();
                        continue;
                    }
                    const int pos = count + __popc(bal & ((1u << lane) - 1u));
                    if (take) { sd[pos] = d2; si[pos] = pidx; }
                    count += __popc(bal);
                    __syncwarp();
                }
                // once the own cell / the face neighbours are in and no threshold exists yet, pay one sort to
                // get one: everything farther than the current k-th best is then rejected without insertion
                if ((ci == 0 || ci == 6 || ci == 18) && !full_once && count >= k) {
                    double T;
                    count = warp_select_compact<false>(sd, si, count, k, lane, T);
                    if (T < Td) {   // never loosen a bound carried over from a failed level
                        Td = T;
                        Ti = 0x7fffffff;
                        thr_incl = true;
                    }
                    full_once = true;
                }
            }
            // final sort of what we have.  With more than k entries, first cut the list to the k best exactly (selection
            // on the distance; if distances tie across the cut the tied entries all stay and the full-size sort decides by
            // index), so that the exact (d2, index) sort runs on half the network
            int sort_n = cap;
            if (count > k) {
                double T;
                count = warp_select_compact<true>(sd, si, count, k, lane, T);
            }
            if (count == k) sort_n = k2;
            for (int t = count + lane; t < sort_n; t += 32) { sd[t] = INF; si[t] = 0x7fffffff; }
            __syncwarp();
            warp_bitonic_sort(sd, si, sort_n, lane);
            const bool have_k = count >= k;
            const double dk2 = have_k ? sd[k - 1] : INF;
            if (top || (have_k && dk2 < margin2)) {
                int* oi = a.out_idx + qw * (long long)k;
                for (int t = lane; t < k; t += 32) {
                    oi[t] = si[t];
                    if (a.out_dist) a.out_dist[qw * (long long)k + t] = sqrt(sd[t]);
                }
                if (a.out_rad && lane == 0) a.out_rad[qw] = sqrt(sd[k - 1]);
                level_hint = l;
                __syncwarp();
                break;
            }
            // failed: carry the k-th key (an upper bound of the true k-th key) to the next level
            if (have_k) {
                Td = sd[k - 1];
                Ti = si[k - 1];
            } else {
                Td = INF;
                Ti = 0x7fffffff;
            }
            thr_incl = true;
            __syncwarp();
        }
    }
}


// ------------------------------------------------------------------------------------------------ ball query
// Fixed-radius neighbourhoods (reference dataset.py:289-291, cKDTree.query_ball_point): every point with
// d2 <= r*r, d2 evaluated exactly like the kNN key.  One warp per query walks the 27-cell block of the finest grid
// level whose guaranteed-complete radius covers r.  Two passes over the same traversal: COUNT, then (after an exclusive
// scan of the counts on the caller's side) FILL at the query's offset -- in traversal order, which is deterministic;
// callers that need a canonical order sort each segment by index.
struct BallArgs {
    const float4* sorted;
    const uint32_t* tab;
    const float* pts;
    const GridParams* gp;
    const int* query;
    long long q_begin, q_count;
    const double* radius;   // per query, or NULL -> r
    double r;
    long long* count;       // COUNT pass
    const long long* offsets;   // FILL pass
    int* out_idx;
};

template <bool FILL>
__global__ void __launch_bounds__(kKnnWarps * 32) ball_kernel(BallArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GridParams gpar = *a.gp;
    const int G = 1 << gpar.bits;
    for (long long qw = (long long)blockIdx.x * kKnnWarps + warp; qw < a.q_count; qw += (long long)gridDim.x * kKnnWarps) {
        const long long qi = a.query ? (long long)a.query[qw] : a.q_begin + qw;
        const float qxf = a.pts[qi * 3], qyf = a.pts[qi * 3 + 1], qzf = a.pts[qi * 3 + 2];
        const double qx = qxf, qy = qyf, qz = qzf;
        const double r = a.radius ? a.radius[qw] : a.r;
        const double r2 = r * r;
        const float ux = cell_coord(qxf, gpar.minx, gpar.inv_cell), uy = cell_coord(qyf, gpar.miny, gpar.inv_cell),
                    uz = cell_coord(qzf, gpar.minz, gpar.inv_cell);
        const int ix = cell_index(ux, G), iy = cell_index(uy, G), iz = cell_index(uz, G);
        // finest level whose block is guaranteed to contain the ball (sides on the domain boundary impose no limit)
        int l = 0;
        for (; l < gpar.bits; ++l) {
            const int Gl = G >> l;
            const float w = (float)(1 << l);
            const float uu[3] = {ux, uy, uz};
            const int cc[3] = {ix >> l, iy >> l, iz >> l};
            float mrg = 3.0e38f;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (cc[t] - 1 > 0) mrg = fminf(mrg, uu[t] - (float)(cc[t] - 1) * w - kCellSlack);
                if (cc[t] + 2 < Gl) mrg = fminf(mrg, (float)(cc[t] + 2) * w - uu[t] - kCellSlack);
            }
            if (mrg >= 3.0e38f) break;
            const double mm = (double)fmaxf(mrg, 0.f) * (double)gpar.cell * (1.0 - 1e-6);
            if (mm > r) break;
        }
        const int Gl = G >> l;
        const int cx = ix >> l, cy = iy >> l, cz = iz >> l;
        uint32_t lo = 0, hi = 0;
        float cmin2 = 0.f;
        if (lane < 27) {
            const int ex = cx + (lane % 3) - 1, ey = cy + ((lane / 3) % 3) - 1, ez = cz + (lane / 9) - 1;
            if (ex >= 0 && ey >= 0 && ez >= 0 && ex < Gl && ey < Gl && ez < Gl) {
                const uint32_t m = morton3(ex, ey, ez);
                const int sh = 3 * l;
                if (sh >= 30) {
                    lo = 0;
                    hi = (uint32_t)gpar.n;
                } else {
                    lo = a.tab[(size_t)m << sh];
                    hi = a.tab[((size_t)m + 1) << sh];
                }
                const float w = (float)(1 << l);
                float dd[3];
                const float uu[3] = {ux, uy, uz};
                const int ee[3] = {ex, ey, ez};
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float c0 = (float)ee[t] * w, c1 = c0 + w;
                    float d = 0.f;
                    if (uu[t] < c0) d = c0 - uu[t] - kCellSlack;
                    else if (uu[t] > c1) d = uu[t] - c1 - kCellSlack;
                    dd[t] = fmaxf(d, 0.f);
                }
                const float dc = gpar.cell * (1.f - 1e-5f);
                cmin2 = (dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]) * dc * dc * (1.f - 1e-5f);
            }
        }
        long long n = 0;
        int* out = FILL ? a.out_idx + a.offsets[qw] : nullptr;
        for (int c = 0; c < 27; ++c) {
            const uint32_t clo = __shfl_sync(0xffffffffu, lo, c), chi = __shfl_sync(0xffffffffu, hi, c);
            const float cm2 = __shfl_sync(0xffffffffu, cmin2, c);
            if (chi <= clo || (double)cm2 > r2) continue;
            for (uint32_t j0 = clo; j0 < chi; j0 += 32) {
                const uint32_t j = j0 + lane;
                bool in = false;
                int pidx = 0;
                if (j < chi) {
                    const float4 p = __ldg(a.sorted + j);
                    const double dx = (double)p.x - qx, dy = (double)p.y - qy, dz = (double)p.z - qz;
                    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    pidx = __float_as_int(p.w);
                    in = d2 <= r2;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (FILL && in) out[n + __popc(bal & ((1u << lane) - 1u))] = pidx;
                n += __popc(bal);
            }
        }
        if (!FILL && lane == 0) a.count[qw] = n;
    }
}

}  // namespace
}  // namespace dfb

struct dfb_grid {
    long long n = 0;
    int bits = 0;
    size_t ncell = 0;
    long long ntile = 0;
    float* pts = nullptr;              // device copy, original order
    float4* sorted = nullptr;          // device, Morton order (x, y, z, original index)
    uint32_t* tab = nullptr;           // device, 8^bits + 1 entries
    uint32_t* code = nullptr;          // device, per-point Morton code (build scratch)
    uint32_t* sums = nullptr;          // device, scan scratch
    unsigned* bb = nullptr;            // device, ordered-int bounding box (6)
    dfb::GridParams* params = nullptr; // device
};

namespace dfb {
namespace {
int grid_build(dfb_grid* g, const float* d_points, cudaStream_t st) {
    const long long n = g->n;
    const int bb_blocks = (int)((n + 255) / 256 < (long long)sm_count() * 8 ? (n + 255) / 256 : (long long)sm_count() * 8);
    if (d_points != g->pts)
        DFB_CUDA(cudaMemcpyAsync(g->pts, d_points, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DFB_CUDA(cudaMemsetAsync(g->tab, 0, (g->ncell + 1) * sizeof(uint32_t), st));
    bbox_kernel<<<bb_blocks, 256, 0, st>>>(g->pts, n, g->bb);
    params_kernel<<<1, 1, 0, st>>>(g->bb, g->params, g->bits, n);
    const unsigned nb = (unsigned)((n + 255) / 256);
    count_kernel<<<nb, 256, 0, st>>>(g->pts, g->params, g->code, g->tab);
    scan_tile_sums<<<(unsigned)g->ntile, kScanThreads, 0, st>>>(g->tab + 1, (long long)g->ncell, g->sums);
    scan_sums_serial<<<1, 1024, 0, st>>>(g->sums, g->ntile);
    scan_apply<<<(unsigned)g->ntile, kScanThreads, 0, st>>>(g->tab + 1, (long long)g->ncell, g->sums);
    scatter_kernel<<<nb, 256, 0, st>>>(g->pts, n, g->code, g->tab, g->sorted);
    note_launch(7);
    return check_cuda(cudaGetLastError(), "grid build kernels", __FILE__, __LINE__);
}
}  // namespace
}  // namespace dfb

namespace dfb {
const float* grid_points(const dfb_grid* g) { return g->pts; }
long long grid_size(const dfb_grid* g) { return g->n; }
}  // namespace dfb

extern "C" int dfb_grid_create(const float* d_points, int64_t n, dfb_stream stream, dfb_grid** out) {
    using namespace dfb;
    DFB_REQUIRE(out != nullptr, DFB_E_ARG, "dfb_grid_create: out is NULL");
    *out = nullptr;
    DFB_REQUIRE(d_points != nullptr && n >= 1 && n < (1LL << 31), DFB_E_ARG, "dfb_grid_create: bad arguments (n=%lld)", (long long)n);
    int rc = dfb_device_check();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    // finest level: about sqrt(n/8) cells per axis for surface-like clouds, 2^5 .. 2^10
    // (exact powers of four must not jump a level: 2 097 152 points -> 2^9, not 2^10 with its 4 GiB table)
    int bits = (int)ceil(0.5 * log2((double)n / 8.0) - 1e-9);
    if (bits < 5) bits = 5;
    if (bits > kMaxLevelBits) bits = kMaxLevelBits;
    dfb_grid* g = new dfb_grid();
    g->n = (long long)n;
    g->bits = bits;
    g->ncell = (size_t)1 << (3 * bits);
    g->ntile = (long long)((g->ncell + kScanTile - 1) / kScanTile);
    cudaError_t e;
    if ((e = cudaMalloc(&g->pts, (size_t)n * 3 * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc(&g->sorted, (size_t)n * sizeof(float4))) != cudaSuccess ||
        (e = cudaMalloc(&g->tab, (g->ncell + 1) * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->code, (size_t)n * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->sums, (size_t)g->ntile * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->bb, 6 * sizeof(unsigned))) != cudaSuccess ||
        (e = cudaMalloc(&g->params, sizeof(GridParams))) != cudaSuccess) {
        dfb_grid_destroy(g);
        return check_cuda(e, "dfb_grid_create allocations", __FILE__, __LINE__);
    }
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    e = cudaMemcpyAsync(g->bb, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `init` lives on this stack frame
    if (e != cudaSuccess) {
        dfb_grid_destroy(g);
        return check_cuda(e, "dfb_grid_create init", __FILE__, __LINE__);
    }
    rc = grid_build(g, d_points, st);
    if (rc) {
        dfb_grid_destroy(g);
        return rc;
    }
    *out = g;
    return DFB_OK;
}

extern "C" int dfb_grid_update(dfb_grid* g, const float* d_points, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(g != nullptr && d_points != nullptr, DFB_E_ARG, "dfb_grid_update: NULL argument");
    return grid_build(g, d_points, as_stream(stream));
}

extern "C" int dfb_grid_destroy(dfb_grid* g) {
    if (!g) return DFB_OK;
    cudaFree(g->pts);
    cudaFree(g->sorted);
    cudaFree(g->tab);
    cudaFree(g->code);
    cudaFree(g->sums);
    cudaFree(g->bb);
    cudaFree(g->params);
    delete g;
    return DFB_OK;
}

extern "C" int dfb_knn(const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count, int32_t k,
                       int32_t* d_idx, double* d_dist, double* d_rad, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(grid != nullptr, DFB_E_ARG, "dfb_knn: grid is NULL");
    DFB_REQUIRE(k >= 1 && k <= 1024 && (long long)k <= grid->n, DFB_E_ARG,
                "dfb_knn: k=%d out of range (1..min(n=%lld,1024))", k, grid->n);
    DFB_REQUIRE(q_count >= 0 && (q_count == 0 || d_idx != nullptr), DFB_E_ARG, "dfb_knn: bad output arguments");
    DFB_REQUIRE(d_query != nullptr || (q_begin >= 0 && q_begin + q_count <= grid->n), DFB_E_ARG,
                "dfb_knn: query range [%lld,%lld) outside the cloud", (long long)q_begin, (long long)(q_begin + q_count));
    if (q_count == 0) return DFB_OK;
    int kcap = 64;
    while (kcap < 2 * k) kcap <<= 1;
    KnnArgs a;
    a.sorted = grid->sorted; a.tab = grid->tab; a.pts = grid->pts; a.gp = grid->params;
    a.query = d_query; a.q_begin = q_begin; a.q_count = q_count;
    a.k = k; a.kcap = kcap;
    a.out_idx = d_idx; a.out_dist = d_dist; a.out_rad = d_rad;
    const size_t smem = (size_t)kKnnWarps * kcap * (sizeof(double) + sizeof(int));
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&knn_kernel), 100 * 1024)) return rc;
    long long blocks = (q_count + kKnnWarps - 1) / kKnnWarps;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    knn_kernel<<<(unsigned)blocks, kKnnWarps * 32, smem, as_stream(stream)>>>(a);
    note_launch();
    return check_cuda(cudaGetLastError(), "knn_kernel launch", __FILE__, __LINE__);
}

namespace dfb {
namespace {
int ball_launch(bool fill, const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count, const double* d_radius,
                double radius, int64_t* d_count, const int64_t* d_offsets, int32_t* d_idx, dfb_stream stream) {
    DFB_REQUIRE(grid != nullptr, DFB_E_ARG, "dfb_ball_*: grid is NULL");
    DFB_REQUIRE(q_count >= 0, DFB_E_ARG, "dfb_ball_*: negative query count");
    DFB_REQUIRE(d_radius != nullptr || radius >= 0.0, DFB_E_ARG, "dfb_ball_*: negative radius");
    DFB_REQUIRE(d_query != nullptr || (q_begin >= 0 && q_begin + q_count <= grid->n), DFB_E_ARG,
                "dfb_ball_*: query range [%lld,%lld) outside the cloud", (long long)q_begin, (long long)(q_begin + q_count));
    if (q_count == 0) return DFB_OK;
    DFB_REQUIRE(fill ? (d_offsets != nullptr && d_idx != nullptr) : d_count != nullptr, DFB_E_ARG, "dfb_ball_*: NULL output");
    BallArgs a;
    a.sorted = grid->sorted; a.tab = grid->tab; a.pts = grid->pts; a.gp = grid->params;
    a.query = d_query; a.q_begin = q_begin; a.q_count = q_count;
    a.radius = d_radius; a.r = radius;
    a.count = reinterpret_cast<long long*>(d_count);
    a.offsets = reinterpret_cast<const long long*>(d_offsets);
    a.out_idx = d_idx;
    long long blocks = (q_count + kKnnWarps - 1) / kKnnWarps;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (fill) ball_kernel<true><<<(unsigned)blocks, kKnnWarps * 32, 0, as_stream(stream)>>>(a);
    else ball_kernel<false><<<(unsigned)blocks, kKnnWarps * 32, 0, as_stream(stream)>>>(a);
    note_launch();
    return check_cuda(cudaGetLastError(), "ball_kernel launch", __FILE__, __LINE__);
}
}  // namespace
}  // namespace dfb

extern "C" int dfb_ball_count(const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count,
                              const double* d_radius, double radius, int64_t* d_count, dfb_stream stream) {
    return dfb::ball_launch(false, grid, d_query, q_begin, q_count, d_radius, radius, d_count, nullptr, nullptr, stream);
}

extern "C" int dfb_ball_fill(const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count,
                             const double* d_radius, double radius, const int64_t* d_offsets, int32_t* d_idx,
                             dfb_stream stream) {
    return dfb::ball_launch(true, grid, d_query, q_begin, q_count, d_radius, radius, nullptr, d_offsets, d_idx, stream);
}
