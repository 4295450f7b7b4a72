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
//   sub-grids (variable density): a finest cell holding more than kHeavy points gets 1 .. 5 levels of its own below
//           the dense table (8^S sub-cells, S from its population, the table in a pool, found through a hash of the
//           cell's Morton code); its points are counting-sorted a second time by their sub-cell code, so every
//           sub-cell of every sub-level is again one contiguous range.  Queries descend into these "negative" levels
//           the same way they move between the dense ones; a neighbouring cell that is not subdivided (that deep)
//           contributes its finest available ancestor's range once.  Uniform clouds have no heavy cell and never
//           leave the dense levels.
#include "common.cuh"

#include <math.h>

#include <algorithm>

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
    int sub_levels;  // deepest sub-grid below the dense table (0: no heavy cell)
    int n_heavy;     // cells with a sub-grid
    unsigned long long pool_used;   // entries of the sub-table pool handed out
};

// ---- sub-grids of heavy cells
constexpr int kHeavy = 512;          // a finest cell with more points than this is subdivided
constexpr int kMaxSub = 5;           // at most 5 levels (32768 sub-cells)
// levels for a cell of c points: about 32 points per occupied sub-cell of a surface-like cloud (4^S occupied sub-cells)
__host__ __device__ __forceinline__ int sub_levels_for(uint32_t c) {
    int S = 1;
    while (S < kMaxSub && (32u << (2 * S)) < c) ++S;
    return S;
}
// hash of cell code -> (S, pool offset): open addressing, entry = (code + 1) << 34 | S << 31 | offset (0 = empty)
__device__ __forceinline__ uint32_t sub_hash(uint32_t m, uint32_t mask) { return (m * 2654435761u) >> 7 & mask; }
__device__ __forceinline__ bool sub_find(const unsigned long long* __restrict__ ht, uint32_t mask, uint32_t m, int& S, uint32_t& off) {
    uint32_t h = sub_hash(m, mask);
    for (;;) {
        const unsigned long long e = __ldg(ht + h);
        if (e == 0ull) return false;
        if ((uint32_t)(e >> 34) == m + 1u) {
            S = (int)((e >> 31) & 7ull);
            off = (uint32_t)(e & 0x7fffffffull);
            return true;
        }
        h = (h + 1u) & mask;
    }
}

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
    gp->sub_levels = 0;
    gp->n_heavy = 0;
    gp->pool_used = 0ull;
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

// After the build the scratch array is free: it keeps the cloud as float4 in ORIGINAL order (grid_points4), so that
// the patch gather of stage 2 costs one 16-byte load per neighbour instead of three 4-byte ones.
__global__ void points4_kernel(const float* __restrict__ pts, long long n, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = make_float4(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], 0.f);
}

// ---- second sort: the points of heavy cells by sub-cell code
struct SubBuild {
    const float4* in;             // sorted by finest dense cell (scatter_kernel)
    float4* out;                  // same, heavy cells ordered by sub-cell code as well
    long long n;
    const uint32_t* code;         // dense Morton code per ORIGINAL point index
    const uint32_t* tab;
    GridParams* gp;
    unsigned long long* ht;
    uint32_t ht_mask;
    uint32_t* pool;
    unsigned long long pool_cap;
    uint32_t* heavy;              // list of heavy cell codes
    uint32_t heavy_cap;
};
// sub-cell code (3 S bits) of a point inside its dense cell
__device__ __forceinline__ uint32_t sub_code(const GridParams& g, float x, float y, float z, int S) {
    const int Gs = (1 << g.bits) << S;
    const float sc = (float)(1 << S);
    const uint32_t msk = (1u << S) - 1u;
    const int fx = cell_index(cell_coord(x, g.minx, g.inv_cell) * sc, Gs), fy = cell_index(cell_coord(y, g.miny, g.inv_cell) * sc, Gs),
              fz = cell_index(cell_coord(z, g.minz, g.inv_cell) * sc, Gs);
    return morton3((uint32_t)fx & msk, (uint32_t)fy & msk, (uint32_t)fz & msk);
}
// the first point of every heavy cell registers the cell: pool space, hash entry, list entry
__global__ void sub_mark_kernel(SubBuild b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const uint32_t m = b.code[__float_as_int(b.in[i].w)];
    const uint32_t lo = b.tab[m];
    if ((long long)lo != i) return;
    const uint32_t c = b.tab[m + 1] - lo;
    if (c <= (uint32_t)kHeavy) return;
    const int S = sub_levels_for(c);
    const unsigned long long cells = (1ull << (3 * S)) + 1ull;
    const unsigned long long off = atomicAdd(&b.gp->pool_used, cells);
    if (off + cells > b.pool_cap || off + cells > 0x7fffffffull) return;   // no room: the cell stays undivided
    const uint32_t slot = (uint32_t)atomicAdd(&b.gp->n_heavy, 1);
    if (slot >= b.heavy_cap) return;
    b.heavy[slot] = m;
    atomicMax(&b.gp->sub_levels, S);
    const unsigned long long e = ((unsigned long long)(m + 1u) << 34) | ((unsigned long long)S << 31) | off;
    uint32_t h = sub_hash(m, b.ht_mask);
    while (atomicCAS(b.ht + h, 0ull, e) != 0ull) h = (h + 1u) & b.ht_mask;
}
__global__ void sub_count_kernel(SubBuild b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const float4 p = b.in[i];
    const uint32_t m = b.code[__float_as_int(p.w)];
    if (b.tab[m + 1] - b.tab[m] <= (uint32_t)kHeavy) return;
    int S;
    uint32_t off;
    if (!sub_find(b.ht, b.ht_mask, m, S, off)) return;
    atomicAdd(b.pool + off + 1 + sub_code(*b.gp, p.x, p.y, p.z, S), 1u);
}
// one CTA per heavy cell: exclusive scan of its sub-cell counts, based at the cell's first point
__global__ void __launch_bounds__(256) sub_scan_kernel(SubBuild b) {
    __shared__ uint32_t s_part[256];
    const int nh = min(b.gp->n_heavy, (int)b.heavy_cap);
    for (int hi = blockIdx.x; hi < nh; hi += gridDim.x) {
        const uint32_t m = b.heavy[hi];
        int S;
        uint32_t off;
        if (!sub_find(b.ht, b.ht_mask, m, S, off)) continue;
        const int cells = 1 << (3 * S);
        uint32_t* t = b.pool + off + 1;
        const int per = (cells + 255) / 256;
        const int j0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (int j = j0; j < min(cells, j0 + per); ++j) sum += t[j];
        s_part[threadIdx.x] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = b.tab[m];
            for (int q = 0; q < 256; ++q) { const uint32_t v = s_part[q]; s_part[q] = run; run += v; }
            b.pool[off] = b.tab[m];
        }
        __syncthreads();
        uint32_t run = s_part[threadIdx.x];
        for (int j = j0; j < min(cells, j0 + per); ++j) { const uint32_t v = t[j]; t[j] = run; run += v; }
        __syncthreads();
    }
}
__global__ void sub_scatter_kernel(SubBuild b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    const float4 p = b.in[i];
    const uint32_t m = b.code[__float_as_int(p.w)];
    long long pos = i;
    if (b.tab[m + 1] - b.tab[m] > (uint32_t)kHeavy) {
        int S;
        uint32_t off;
        if (sub_find(b.ht, b.ht_mask, m, S, off)) pos = atomicAdd(b.pool + off + 1 + sub_code(*b.gp, p.x, p.y, p.z, S), 1u);
    }
    b.out[pos] = p;
}

// Indices of the points with q_begin <= index < q_end in MORTON order (the order of the sorted copy): an order-preserving
// stream compaction of the sorted array in three launches (per-block counts, one-block exclusive scan, scatter).
// Spatially ordered query lists are what makes consecutive queries share their candidate cells (dfb_knn).
constexpr int kOrderBlock = 1024;
__global__ void __launch_bounds__(kOrderBlock) order_count_kernel(const float4* __restrict__ sorted, long long n, int q0, int q1,
                                                                    uint32_t* __restrict__ counts) {
    __shared__ uint32_t s_w[kOrderBlock / 32];
    const long long i = (long long)blockIdx.x * kOrderBlock + threadIdx.x;
    bool in = false;
    if (i < n) {
        const int id = __float_as_int(sorted[i].w);
        in = id >= q0 && id < q1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = s_w[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) counts[blockIdx.x] = v;
    }
}
__global__ void __launch_bounds__(kOrderBlock) order_scatter_kernel(const float4* __restrict__ sorted, long long n, int q0, int q1,
                                                                      const uint32_t* __restrict__ base, int* __restrict__ out) {
    __shared__ uint32_t s_w[kOrderBlock / 32];
    const long long i = (long long)blockIdx.x * kOrderBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    bool in = false;
    int id = 0;
    if (i < n) {
        id = __float_as_int(sorted[i].w);
        in = id >= q0 && id < q1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_w[w] = __popc(bal);
    __syncthreads();
    if (threadIdx.x < 32) {   // exclusive scan of the 32 warp counts
        const uint32_t v = s_w[threadIdx.x];
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (threadIdx.x >= o) x += y;
        }
        s_w[threadIdx.x] = x - v;
    }
    __syncthreads();
    if (in) out[base[blockIdx.x] + s_w[w] + __popc(bal & ((1u << lane) - 1u))] = id;
}

// ---------------------------------------------------------------------------------------- query

// Warp reductions whose result the compiler KNOWS to be warp-uniform (REDUX, one instruction): a butterfly of xor-shuffles
// yields the same value in every lane too, but the compiler cannot prove it, so every branch on such a value makes the
// code that follows "potentially divergent" and each later shuffle / vote is wrapped in a WARPSYNC.COLLECTIVE sequence
// (5 instructions instead of 1; 190 sites in this kernel).
__device__ __forceinline__ int warp_sum_uniform(int v) { return (int)__reduce_add_sync(0xffffffffu, (unsigned)v); }
// minimum / maximum of non-negative 64-bit patterns
__device__ __forceinline__ long long warp_min_uniform(long long v) {
    const unsigned h = __reduce_min_sync(0xffffffffu, (unsigned)(v >> 32));
    const unsigned l = __reduce_min_sync(0xffffffffu, (unsigned)(v >> 32) == h ? (unsigned)v : 0xffffffffu);
    return (long long)(((unsigned long long)h << 32) | l);
}
__device__ __forceinline__ long long warp_max_uniform(long long v) {
    const unsigned h = __reduce_max_sync(0xffffffffu, (unsigned)(v >> 32));
    const unsigned l = __reduce_max_sync(0xffffffffu, (unsigned)(v >> 32) == h ? (unsigned)v : 0u);
    return (long long)(((unsigned long long)h << 32) | l);
}

// The kernel body is far larger than the instruction cache (6 300 SASS instructions = 100 KB at k = 256; ncu: 17 % of the
// stall samples are instruction fetches), and the selection and the shared-memory sorting network were inlined at three
// and two call sites: as real functions they exist once (DFB_KNN_INLINE_HELPERS restores the inlining for A/B builds).
#ifdef DFB_KNN_INLINE_HELPERS
#define DFB_KNN_HELPER __device__ __forceinline__
#else
#define DFB_KNN_HELPER __device__ __noinline__
#endif

struct Key {
    double d2;
    int idx;
};
__device__ __forceinline__ bool key_less(double a, int ai, double b, int bi) { return a < b || (a == b && ai < bi); }

// Ascending bitonic sort of cap (power of two) entries held in shared memory by one warp.
DFB_KNN_HELPER void warp_bitonic_sort(double* sd, int* si, int cap, int lane) {
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

// Tighten the list without sorting it: find a bound T with at least k (and at most a few more) of the count entries
// <= T, then compact those to the front.  T is an UPPER bound of the k-th smallest key seen so far, which is all the
// pruning needs (every true neighbour has d2 <= T; ties are kept, the final sort orders them).
// The search keeps a bracket (a, b] of BIT PATTERNS (monotone in the non-negative distances) with #(<= a) < k <= #(<= b)
// and places the next probe by linear interpolation of the counts in VALUE space: squared distances to the points of a
// surface or a volume are close to uniformly / polynomially distributed, so two or three probes usually land inside the
// [k, k + 16] window where 10+ bisection steps were needed (the bisection loop was half of the kernel's instructions).
// A probe that fails to shrink the bracket from the same side twice falls back to the midpoint of the bit patterns, and
// as an integer search the loop ends at the exact k-th value at the latest.
// Returns the new count (>= k; larger than k + 16 only for exact ties).  EXACT: run the search until exactly k entries
// are <= T (or the bracket is one bit pattern wide: distances tie across the cut and the tied entries all stay).
// One function with EXACT as a run-time flag: two instantiations were 1 300 instructions of a kernel that does not fit
// the instruction cache.
DFB_KNN_HELPER int warp_select_compact(double* sd, int* si, int count, int k, int lane, double& T, bool EXACT) {
    long long lo = 0x7ff0000000000000LL, hi = 0;
    for (int t = lane; t < count; t += 32) {
        const long long v = __double_as_longlong(sd[t]);
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = warp_min_uniform(lo);
    hi = warp_max_uniform(hi);
    // bracket: #(<= a) = ca < k <= cb = #(<= b)
    long long a = lo - 1, b = hi;
    int ca = 0, cb = count;
    const int want = EXACT ? k : min(k + 8, count);
    int same_side = 0;   // +n: the last n probes moved b, -n: moved a
    if (!(EXACT ? cb == k : cb <= k + 16)) {
        for (int iter = 0; iter < 96 && b - a > 1; ++iter) {
            long long mid;
            if (same_side >= 2 || same_side <= -2 || iter >= 12) {
                mid = a + ((b - a) >> 1);
                same_side = 0;
            } else {
                const double av = __longlong_as_double(max(a, 0LL)), bv = __longlong_as_double(b);
                const double g = av + (bv - av) * ((double)(want - ca) / (double)(cb - ca));
                mid = __double_as_longlong(g);
                mid = max(a + 1, min(b - 1, mid));
            }
            int c = 0;
            for (int t = lane; t < count; t += 32) c += (__double_as_longlong(sd[t]) <= mid) ? 1 : 0;
            c = warp_sum_uniform(c);
            if (c >= k) {
                b = mid;
                cb = c;
                same_side = same_side > 0 ? same_side + 1 : 1;
                if (EXACT ? c == k : c <= k + 16) break;
            } else {
                a = mid;
                ca = c;
                same_side = same_side < 0 ? same_side - 1 : -1;
            }
        }
    }
    hi = b;
    T = __longlong_as_double(hi);
    if (cb == count) return count;   // nothing to drop
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

// Ascending sort by (d2, index) of 32 * E entries held in registers, E consecutive positions per lane (position
// lane * E + r): the strides below E are compare-exchanges between registers of one lane, the others one xor-shuffle of
// the three key words.  d2 travels as its bit pattern (monotone for non-negative doubles) and a comparison is one 96-bit
// integer compare of (d2, index).  Replaces the shared-memory network for the final exactly-k sort (a quarter of the
// kernel's instructions and the source of its bank conflicts).  The (size, stride) loops stay ROLLED -- one inter-lane
// stage body plus log2(E) intra-lane bodies, a few hundred instructions: the fully unrolled network (19 K instructions)
// thrashed the instruction cache (50 % of the stall samples were no-instruction stalls).
__device__ __forceinline__ bool key_less_bits(long long a, int ai, long long b, int bi) {
    const unsigned __int128 ka = ((unsigned __int128)(unsigned long long)a << 32) | (unsigned)ai;
    const unsigned __int128 kb = ((unsigned __int128)(unsigned long long)b << 32) | (unsigned)bi;
    return ka < kb;
}

template <int E>
__device__ __forceinline__ void warp_sort_regs(long long (&d)[E], int (&id)[E], int lane) {
    constexpr int N = 32 * E;
    const int base = lane * E;
#pragma unroll 1
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride >= E; stride >>= 1) {
            const int m = stride / E;
            // every position of this lane has the same (e & stride) and (e & size) bits
            const bool keep_min = ((base & stride) == 0) == ((base & size) == 0);
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const long long od = __shfl_xor_sync(0xffffffffu, d[r], m);
                const int oi = __shfl_xor_sync(0xffffffffu, id[r], m);
                // keys are distinct (equal keys are identical padding entries, for which the choice does not matter)
                if (key_less_bits(od, oi, d[r], id[r]) == keep_min) { d[r] = od; id[r] = oi; }
            }
        }
#pragma unroll
        for (int st = E >> 1; st > 0; st >>= 1) {
            if (st <= (size >> 1)) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & st) == 0) {
                        const int r2 = r + st;
                        const bool asc = ((base + r) & size) == 0;
                        if (key_less_bits(d[r2], id[r2], d[r], id[r]) == asc) {
                            const long long td = d[r]; d[r] = d[r2]; d[r2] = td;
                            const int ti = id[r]; id[r] = id[r2]; id[r2] = ti;
                        }
                    }
                }
            }
        }
    }
}

// Final step of a query whose list holds exactly k entries: register sort, then either the outputs (the k-th distance
// lies strictly inside the block's guaranteed margin, or the block is the whole cloud) or the k-th key for the retry one
// level up.  Deliberately NOT inlined: inside the query loop the compiler multiplies the sort's live ranges across its
// cloned loop bodies (227 registers at E = 8); as a function it needs 40.
struct SortResult {
    long long kd;   // bit pattern of the k-th squared distance
    int kid;        // its index
    int accepted;
};
template <int E>
__device__ __forceinline__ SortResult sort_emit(const double* sd, const int* si, int k, int lane, bool top, double margin2, int* oi,
                                             double* odist, double* orad) {
    long long kd[E];
    int kid[E];
    // the list is unordered, so the conflict-free striped read may land anywhere
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int t = r * 32 + lane;
        kd[r] = t < k ? __double_as_longlong(sd[t]) : 0x7ff0000000000000LL;
        kid[r] = t < k ? si[t] : 0x7fffffff;
    }
    warp_sort_regs<E>(kd, kid, lane);
    SortResult res;
    res.kd = 0;
    res.kid = 0;
    // position k - 1 sits in lane (k - 1) / E: the lane's last entry with position < k.  (Written as a monotone select
    // chain: picking register (k - 1) % E by comparison makes the compiler index the arrays dynamically, i.e. move them
    // to local memory.)
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const bool below = lane * E + r < k;
        res.kd = below ? kd[r] : res.kd;
        res.kid = below ? kid[r] : res.kid;
    }
    res.kd = __shfl_sync(0xffffffffu, res.kd, (k - 1) / E);
    res.kid = __shfl_sync(0xffffffffu, res.kid, (k - 1) / E);
    const double dk2 = __longlong_as_double(res.kd);
    res.accepted = (top || dk2 < margin2) ? 1 : 0;
    if (res.accepted) {
        if ((k % E) == 0 && E >= 4) {   // every lane holds a whole, 16-byte aligned run
            if (lane * E < k) {
#pragma unroll
                for (int r = 0; r + 3 < E; r += 4)
                    *reinterpret_cast<int4*>(oi + lane * E + r) = make_int4(kid[r], kid[r + 1], kid[r + 2], kid[r + 3]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < E; ++r)
                if (lane * E + r < k) oi[lane * E + r] = kid[r];
        }
        if (odist) {
#pragma unroll
            for (int r = 0; r < E; ++r)
                if (lane * E + r < k) odist[lane * E + r] = sqrt(__longlong_as_double(kd[r]));
        }
        if (orad && lane == 0) *orad = sqrt(dk2);
    }
    return res;
}

// The same final step without a sorting network.  The k squared distances of a neighbourhood spread almost evenly over
// [0, d_k^2] (a ball on a surface holds points in proportion to r^2), so floor(d2 * k / d_k^2) is a nearly perfect,
// MONOTONE hash into k buckets: a counting sort by bucket (shared-memory atomics, one warp scan) leaves every entry
// within a few places of its final position, and the exact (d2, index) rank inside a bucket is a loop over its one or
// two members.  About 400 warp instructions where the 36-stage bitonic network of 256 entries with 96-bit keys took
// 4.4 k (ncu: 36 % of the kernel's instructions).  Exact for any input: the bucket function is monotone in d2 (fp64
// product, truncation and clamp all are), ties and crowded buckets are ranked by full key comparison; a neighbourhood
// that piles more than kBucketMax entries into one bucket (a cluster plus a few far points, a lattice shell) takes the
// kernel's general shared-memory sorting network instead (return code 2).  The acceptance test needs only the largest key, so a level that fails it is not sorted at
// all.  The histogram / bucket ends take 32 E ints of shared memory per warp next to the list; the bucket-ordered copy goes
// to the list's upper halves sd[k, 2k) / si[k, 2k), free once the list has been cut to exactly k entries.
constexpr int kBucketMax = 12;
#ifndef DFB_KNN_BU
#define DFB_KNN_BU 2
#endif
constexpr int kBucketUnroll = DFB_KNN_BU;
__device__ __forceinline__ int bucket_of(double d2, double scale, int kb) { return min(kb, (int)(d2 * scale)); }
template <int E>
__device__ __noinline__ SortResult bucket_emit(double* sd, int* si, int* h, int k, int lane, bool top, double margin2, int* oi,
                                               double* odist, double* orad) {
    long long hi = 0;
#pragma unroll (kBucketUnroll)
    for (int t = lane; t < k; t += 32) hi = max(hi, __double_as_longlong(sd[t]));
    hi = warp_max_uniform(hi);
    SortResult res;
    res.kd = hi;
    res.kid = 0;
    const double dk2 = __longlong_as_double(hi);
    res.accepted = (top || dk2 < margin2) ? 1 : 0;
    if (!res.accepted) {   // the retry one level up starts from the largest key
        int mi = -1;
        for (int t = lane; t < k; t += 32) mi = __double_as_longlong(sd[t]) == hi ? max(mi, si[t]) : mi;
        res.kid = __reduce_max_sync(0xffffffffu, mi);
        return res;
    }
    for (int t = lane; t < 32 * E; t += 32) h[t] = 0;
    __syncwarp();
    const double scale = hi > 0 ? (double)k / dk2 : 0.0;
    const int kb = k - 1;
#pragma unroll (kBucketUnroll)
    for (int t = lane; t < k; t += 32) atomicAdd(&h[bucket_of(sd[t], scale, kb)], 1);
    __syncwarp();
    // exclusive scan of the counts, E consecutive buckets per lane: h[b] = start of bucket b; the placement pass then
    // advances every h[b] to the END of its bucket
    int c[E];
    int sum = 0, maxc = 0;
#pragma unroll
    for (int j = 0; j < E; ++j) {
        c[j] = h[lane * E + j];
        maxc = max(maxc, c[j]);
        sum += c[j];
    }
    maxc = __reduce_max_sync(0xffffffffu, maxc);
    if (maxc > kBucketMax) {   // the list itself is untouched so far: the caller's shared-memory network sorts it
        res.accepted = 2;
        return res;
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int run = incl - sum;
#pragma unroll
    for (int j = 0; j < E; ++j) {
        h[lane * E + j] = run;
        run += c[j];
    }
    __syncwarp();
    // placement: the list copied in bucket order into the free upper halves sd[k, 2k), si[k, 2k) (order inside a bucket
    // arbitrary)
    double* td = sd + k;
    int* ti = si + k;
#pragma unroll (kBucketUnroll)
    for (int t = lane; t < k; t += 32) {
        const double d = sd[t];
        const int pos = atomicAdd(&h[bucket_of(d, scale, kb)], 1);
        td[pos] = d;
        ti[pos] = si[t];
    }
    __syncwarp();
    // exact rank inside the bucket, then straight to the output row (ranks follow the positions closely, so the stores
    // of a warp stay nearly contiguous)
#pragma unroll 2
    for (int p = lane; p < k; p += 32) {
        const double dd = td[p];
        const long long d = __double_as_longlong(dd);
        const int id = ti[p];
        const int b = bucket_of(dd, scale, kb);
        const int s = b > 0 ? h[b - 1] : 0, e = h[b];
        int rank = p;
        if (e - s > 1) {
            rank = s;
            for (int j = s; j < e; ++j) {
                const long long dj = __double_as_longlong(td[j]);
                rank += (dj < d || (dj == d && ti[j] < id)) ? 1 : 0;
            }
        }
        oi[rank] = id;
        if (odist) odist[rank] = sqrt(dd);
    }
    if (orad && lane == 0) *orad = sqrt(dk2);
    return res;
}

struct KnnArgs {
    const float4* sorted;
    const uint32_t* tab;
    const unsigned long long* ht;   // sub-grid hash (cell code -> levels, pool offset)
    uint32_t ht_mask;
    const uint32_t* pool;           // sub-grid tables
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

// Range of the sorted array and conservative squared distance (world units) of cell (ex, ey, ez) of level l, one of the
// 27 cells of the block whose first cell is (bx, by, bz) (may be -1).  l >= 0: dense table.  l < 0: sub-level -l of the
// finest cells; where the parent cell is not subdivided that deep, the finest available ancestor stands in for all of the
// block's cells it contains and is returned by the first of them only (the others return an empty range).
#ifdef DFB_KNN_BLOCKCELL_NOINLINE
#define DFB_KNN_BLOCKCELL __device__ __noinline__
#else
#define DFB_KNN_BLOCKCELL __device__ __forceinline__
#endif
// While a radius guess from the neighbouring query is active the bound is tight already (1.5 x its k-th d2): the early
// selection after the own cell / the face neighbours would re-derive what the guess gives (-DDFB_KNN_SELECT_GUESSED
// restores it for A/B builds).
#ifdef DFB_KNN_SELECT_GUESSED
#define DFB_KNN_SKIP_GUESSED
#else
#define DFB_KNN_SKIP_GUESSED !guessed &&
#endif
#ifdef DFB_KNN_EXPECT
#define DFB_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define DFB_UNLIKELY(x) (x)
#endif
DFB_KNN_BLOCKCELL void block_cell(const KnnArgs& a, const GridParams& gpar, int G, int l, int ex, int ey, int ez, int bx,
                                           int by, int bz, float ux, float uy, float uz, uint32_t& lo, uint32_t& hi, float& cmin2) {
    lo = hi = 0u;
    cmin2 = 0.f;
    const int Gl = l >= 0 ? (G >> l) : (G << -l);
    if (ex < 0 || ey < 0 || ez < 0 || ex >= Gl || ey >= Gl || ez >= Gl) return;
    float w;            // box edge in finest-dense-cell units
    int ee[3] = {ex, ey, ez};
    if (l >= 0) {
        const uint32_t m = morton3(ex, ey, ez);
        const int sh = 3 * l;
        if (sh >= 30) {  // single top-level cell
            lo = 0;
            hi = (uint32_t)gpar.n;
        } else {
            lo = a.tab[(size_t)m << sh];
            hi = a.tab[((size_t)m + 1) << sh];
        }
        w = (float)(1 << l);
    } else {
        const int sl = -l;
        const uint32_t m = morton3(ex >> sl, ey >> sl, ez >> sl);
        int S = 0;
        uint32_t off = 0;
        const bool found = sub_find(a.ht, a.ht_mask, m, S, off);
        const int d = found ? min(S, sl) : 0;    // depth of the finest available ancestor (sl: the cell itself)
        const int r = sl - d;
        if (r > 0) {
            const int bb[3] = {max(bx, 0), max(by, 0), max(bz, 0)};
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (ee[t] != bb[t] && ((ee[t] - 1) >> r) == (ee[t] >> r)) return;   // an earlier cell of the block owns the ancestor
                ee[t] >>= r;
            }
        }
        if (d == 0) {
            lo = a.tab[m];
            hi = a.tab[m + 1];
        } else {
            const uint32_t msk = (1u << d) - 1u;
            const uint32_t sm = morton3((uint32_t)ee[0] & msk, (uint32_t)ee[1] & msk, (uint32_t)ee[2] & msk);
            const int sh = 3 * (S - d);
            lo = a.pool[off + (sm << sh)];
            hi = a.pool[off + ((sm + 1u) << sh)];
        }
        w = 1.f / (float)(1 << d);
    }
    const float uu[3] = {ux, uy, uz};
    float dd[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const float c0 = (float)ee[t] * w, c1 = c0 + w;
        float dv = 0.f;
        if (uu[t] < c0) dv = c0 - uu[t] - kCellSlack;
        else if (uu[t] > c1) dv = uu[t] - c1 - kCellSlack;
        dd[t] = fmaxf(dv, 0.f);
    }
    const float dc = gpar.cell * (1.f - 1e-5f);
    cmin2 = (dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]) * dc * dc * (1.f - 1e-5f);
}

// E = register entries per lane of the final sort (32 * E = the power of two >= k); 0: shared-memory network only
// Register budget: 6 CTAs per SM (80 registers, 24 warps) up to k = 256 -- the kernel is bound by per-warp latency, and
// since the final step became a counting sort the extra warps outweigh the ~200 bytes of spills (measured +10 .. +14 %
// over 128 registers, profiles/r2_knn_bucket_ab.log; with the sorting network 96 registers lost 3 %); k = 512 fits only
// three CTAs of shared memory and keeps 128 registers.  DFB_KNN_MINB overrides for A/B builds.
#ifndef DFB_KNN_MINB
#define DFB_KNN_MINB 6
#endif
#define DFB_KNN_BOUNDS __launch_bounds__(kKnnWarps * 32, (E >= 16 ? 4 : DFB_KNN_MINB))
// Final step of a query: counting sort by distance bucket from k = 256 up (E >= 8), the register sorting network below
// that (measured, profiles/r2_knn_bucket_ab.log); -DDFB_KNN_BUCKET_MIN_E=99 / =1 force one of them for A/B builds.
#ifndef DFB_KNN_BUCKET_MIN_E
#define DFB_KNN_BUCKET_MIN_E 8
#endif
template <int E>
__device__ __forceinline__ SortResult final_emit(double* sd, int* si, int* h, int k, int lane, bool top, double margin2, int* oi,
                                                 double* odist, double* orad) {
    if constexpr (E >= DFB_KNN_BUCKET_MIN_E) return bucket_emit<E>(sd, si, h, k, lane, top, margin2, oi, odist, orad);
    else return sort_emit<E>(sd, si, k, lane, top, margin2, oi, odist, orad);
}
template <int E>
__global__ void DFB_KNN_BOUNDS knn_kernel(KnnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the warp index through a broadcast: tells the compiler that it (and every loop bound derived from it) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    double* sd = reinterpret_cast<double*>(smem_raw) + (size_t)warp * a.kcap;
    int* si = reinterpret_cast<int*>(smem_raw + (size_t)kKnnWarps * a.kcap * sizeof(double)) + (size_t)warp * a.kcap;
    int* sh = reinterpret_cast<int*>(smem_raw + (size_t)kKnnWarps * a.kcap * (sizeof(double) + sizeof(int))) + warp * (32 * E);
    const GridParams gpar = *a.gp;
    const int G = 1 << gpar.bits;
    const int SM = gpar.sub_levels;           // levels below the dense table (0 for clouds without heavy cells)
    const float fsc = (float)(1 << SM);
    const int k = a.k, cap = a.kcap;
    int k2 = 32;               // size of the sorting network for exactly k entries
    while (k2 < k) k2 <<= 1;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    int level_hint = 0;   // the level that answered the warp's previous query: the search starts there, first DOWN while
                          // the finer block still holds enough points (a query that jumps from a sparse to a dense
                          // region must not scan a coarse block), then up as usual.  Any level passing the margin test
                          // is exact.

    // Each warp takes a contiguous run of queries: consecutive queries of a spatially ordered list share their block of
    // cells, so the level hint and the radius hint below come from a neighbour.
    double prev_k2 = -1.0;   // k-th squared distance of the warp's previous query (< 0: none yet)
    int pcx = 0, pcy = 0, pcz = 0;   // its cell at level_hint
    const long long n_warps = (long long)gridDim.x * kKnnWarps;
    const long long per_warp = (a.q_count + n_warps - 1) / n_warps;
    const long long q_first = ((long long)blockIdx.x * kKnnWarps + warp) * per_warp;
    const long long q_last = min(a.q_count, q_first + per_warp);
    for (long long qw = q_first; qw < q_last; ++qw) {
        const long long qi = a.query ? (long long)a.query[qw] : a.q_begin + qw;
        const float qxf = a.pts[qi * 3], qyf = a.pts[qi * 3 + 1], qzf = a.pts[qi * 3 + 2];
        const double qx = qxf, qy = qyf, qz = qzf;
        const float ux = cell_coord(qxf, gpar.minx, gpar.inv_cell), uy = cell_coord(qyf, gpar.miny, gpar.inv_cell),
                    uz = cell_coord(qzf, gpar.minz, gpar.inv_cell);
        // cell coordinates at the deepest sub-level; the cell of level l (-SM <= l <= bits) is (fx, fy, fz) >> (l + SM)
        const int fx = cell_index(ux * fsc, G << SM), fy = cell_index(uy * fsc, G << SM), fz = cell_index(uz * fsc, G << SM);

        // threshold: candidates must satisfy key <= T when thr_incl (carried over from a failed
        // level) or key < T otherwise (T = current k-th best of the list)
        // Radius hint: the previous query's k-th distance (x 1.5 in d2) as a first inclusive threshold.  It is a GUESS,
        // not a bound: if fewer than k points of the block pass it, the level is scanned again without it.  When k or
        // more pass, the k nearest points of the block are among them, so everything downstream is unchanged.  A
        // good guess rejects most candidates before insertion, so the list never fills and no intermediate
        // selection runs.
        double Td = INF;
        int Ti = 0x7fffffff;
        bool thr_incl = true;
        bool guessed = false;
        // only from a neighbour (same or adjacent cell of the level that answered it): in a spatially ordered query list
        // that is almost every query, in a random one almost none -- there the hint would be the radius of an unrelated
        // region and a wrong guess costs a second scan
        if (prev_k2 > 0.0 && abs((fx >> (level_hint + SM)) - pcx) <= 1 && abs((fy >> (level_hint + SM)) - pcy) <= 1 &&
            abs((fz >> (level_hint + SM)) - pcz) <= 1) {
            Td = prev_k2 * 1.5;
            guessed = true;
        }

        const long long need = min((long long)3 * k, gpar.n);
        int l0 = level_hint;
        while (l0 > -SM) {   // population of the 27-cell block one level finer
            const int lf = l0 - 1;
            const int bx = (fx >> (lf + SM)) - 1, by = (fy >> (lf + SM)) - 1, bz = (fz >> (lf + SM)) - 1;
            unsigned cnt = 0;
            if (lane < 27) {
                uint32_t flo, fhi;
                float fc;
                block_cell(a, gpar, G, lf, bx + (lane % 3), by + ((lane / 3) % 3), bz + (lane / 9), bx, by, bz, ux, uy, uz, flo, fhi, fc);
                cnt = fhi - flo;
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if ((long long)cnt < need) break;
            l0 = lf;
        }
        for (int l = l0; l <= gpar.bits; ++l) {
            const int Gl = l >= 0 ? (G >> l) : (G << -l);
            const int cx = fx >> (l + SM), cy = fy >> (l + SM), cz = fz >> (l + SM);
            // 27 neighbour cells, one per lane
            uint32_t lo = 0, hi = 0;
            float cmin2 = 0.f;  // conservative squared distance from the query to the cell (world units)
            if (lane < 27)
                block_cell(a, gpar, G, l, cx + (lane % 3) - 1, cy + ((lane / 3) % 3) - 1, cz + (lane / 9) - 1, cx - 1, cy - 1, cz - 1,
                           ux, uy, uz, lo, hi, cmin2);
            unsigned total = hi - lo;
            total = __reduce_add_sync(0xffffffffu, total);
            const bool top = (l == gpar.bits);
            if (!top && (long long)total < need) continue;

            // guaranteed-complete radius of the block (world units), strict
            double margin2 = INF;
            if (!top) {
                const float w = l >= 0 ? (float)(1 << l) : 1.f / (float)(1 << -l);
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
                // Three loads in flight per lane (a rotating register queue in a ROLLED loop, so the body exists once):
                // the stream comes from L2 (600+ cycles) and with one load per iteration the warp spent most of a
                // query waiting on memory.
                const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t j = clo + lane;
                float4 n0 = j < chi ? __ldg(a.sorted + j) : zero4;
                float4 n1 = j + 32 < chi ? __ldg(a.sorted + j + 32) : zero4;
                float4 n2 = j + 64 < chi ? __ldg(a.sorted + j + 64) : zero4;
#pragma unroll 1
                for (uint32_t j0 = clo; j0 < chi; j0 += 32, j += 32) {
                    const float4 p = n0;
                    n0 = n1;
                    n1 = n2;
                    n2 = j + 96 < chi ? __ldg(a.sorted + j + 96) : zero4;
                    const bool valid = j < chi;
                    const double dx = (double)p.x - qx, dy = (double)p.y - qy, dz = (double)p.z - qz;
                    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    const int pidx = __float_as_int(p.w);
                    bool take = valid && (thr_incl ? !key_less(Td, Ti, d2, pidx) : key_less(d2, pidx, Td, Ti));
                    unsigned bal = __ballot_sync(0xffffffffu, take);
                    if (bal == 0u) continue;
                    if (DFB_UNLIKELY(count + 32 > cap)) {
                        // the list is full: tighten the bound by selection; only if that cannot make room (a pile of
                        // exact ties) pad, sort and keep exactly the k best
                        double T;
                        count = warp_select_compact(sd, si, count, k, lane, T, false);
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
                        take = valid && (thr_incl ? !key_less(Td, Ti, d2, pidx) : key_less(d2, pidx, Td, Ti));
                        bal = __ballot_sync(0xffffffffu, take);
                        if (bal == 0u) continue;
                    }
                    const int pos = count + __popc(bal & ((1u << lane) - 1u));
                    if (take) { sd[pos] = d2; si[pos] = pidx; }
                    count += __popc(bal);
                    __syncwarp();
                }
                // once the own cell / the face neighbours are in and no threshold exists yet, pay one sort to
                // get one: everything farther than the current k-th best is then rejected without insertion
                if ((ci == 0 || ci == 6 || ci == 18) && !full_once && DFB_KNN_SKIP_GUESSED count >= k) {
                    double T;
                    count = warp_select_compact(sd, si, count, k, lane, T, false);
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
            if (guessed) {
                guessed = false;
                if (DFB_UNLIKELY(count < k)) {   // the radius hint was too small here: same level again, without it
                    Td = INF;
                    Ti = 0x7fffffff;
                    thr_incl = true;
                    --l;
                    continue;
                }
            }
            if (count > k) {
                double T;
                count = warp_select_compact(sd, si, count, k, lane, T, true);
            }
            if (E > 0 && count == k) {
                __syncwarp();
                const SortResult sr = final_emit<(E > 0 ? E : 1)>(sd, si, sh, k, lane, top, margin2, a.out_idx + qw * (long long)k,
                                                                 a.out_dist ? a.out_dist + qw * (long long)k : nullptr,
                                                                 a.out_rad ? a.out_rad + qw : nullptr);
                if (sr.accepted == 1) {
                    prev_k2 = __longlong_as_double(sr.kd);
                    level_hint = l;
                    pcx = cx; pcy = cy; pcz = cz;
                    __syncwarp();
                    break;
                }
                if (sr.accepted == 0) {
                    Td = __longlong_as_double(sr.kd);
                    Ti = sr.kid;
                    thr_incl = true;
                    __syncwarp();
                    continue;
                }
                // 2: crowded buckets -- the general path below sorts the (untouched) list
            }
            int sort_n = cap;
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
                prev_k2 = sd[k - 1];
                level_hint = l;
                pcx = cx; pcy = cy; pcz = cz;
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
    // Each warp takes a contiguous run of queries: consecutive queries of a spatially ordered list share their block of
    // cells, so the level hint and the radius hint below come from a neighbour.
    double prev_k2 = -1.0;   // k-th squared distance of the warp's previous query (< 0: none yet)
    int pcx = 0, pcy = 0, pcz = 0;   // its cell at level_hint
    const long long n_warps = (long long)gridDim.x * kKnnWarps;
    const long long per_warp = (a.q_count + n_warps - 1) / n_warps;
    const long long q_first = ((long long)blockIdx.x * kKnnWarps + warp) * per_warp;
    const long long q_last = min(a.q_count, q_first + per_warp);
    for (long long qw = q_first; qw < q_last; ++qw) {
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
    float4* sorted = nullptr;          // device, Morton order (x, y, z, original index), heavy cells by sub-cell code too
    float4* sorted_tmp = nullptr;      // device, build scratch (sorted by finest dense cell only); after the build: the
                                       // cloud as float4 in original order (grid_points4)
    unsigned long long* ht = nullptr;  // device, sub-grid hash
    uint32_t ht_size = 0;
    uint32_t* pool = nullptr;          // device, sub-grid tables
    unsigned long long pool_cap = 0;
    uint32_t* heavy = nullptr;         // device, list of heavy cells
    uint32_t heavy_cap = 0;
    uint32_t* tab = nullptr;           // device, 8^bits + 1 entries
    uint32_t* code = nullptr;          // device, per-point Morton code (build scratch)
    uint32_t* sums = nullptr;          // device, scan scratch
    uint32_t* order_cnt = nullptr;     // device, per-block counts of dfb_grid_morton_order
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
    scatter_kernel<<<nb, 256, 0, st>>>(g->pts, n, g->code, g->tab, g->sorted_tmp);
    // second sort inside heavy cells (sub-grids); clouds without heavy cells pay one copy of the sorted array
    DFB_CUDA(cudaMemsetAsync(g->ht, 0, (size_t)g->ht_size * sizeof(unsigned long long), st));
    DFB_CUDA(cudaMemsetAsync(g->pool, 0, (size_t)g->pool_cap * sizeof(uint32_t), st));
    SubBuild sb{g->sorted_tmp, g->sorted, n, g->code, g->tab, g->params, g->ht, g->ht_size - 1, g->pool, g->pool_cap, g->heavy,
                g->heavy_cap};
    sub_mark_kernel<<<nb, 256, 0, st>>>(sb);
    sub_count_kernel<<<nb, 256, 0, st>>>(sb);
    sub_scan_kernel<<<(unsigned)std::min<long long>((long long)g->heavy_cap, (long long)sm_count() * 8), 256, 0, st>>>(sb);
    sub_scatter_kernel<<<nb, 256, 0, st>>>(sb);
    points4_kernel<<<nb, 256, 0, st>>>(g->pts, n, g->sorted_tmp);
    note_launch(12);
    return check_cuda(cudaGetLastError(), "grid build kernels", __FILE__, __LINE__);
}
}  // namespace
}  // namespace dfb

namespace dfb {
const float* grid_points(const dfb_grid* g) { return g->pts; }
const float4* grid_points4(const dfb_grid* g) { return g->sorted_tmp; }
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
    // sub-grids: at most n / kHeavy heavy cells; a cell of c points takes 8^S + 1 <= 4 c + 1 table entries, usually about
    // c: room for 2 n (a cell that finds the pool full stays undivided, which costs speed, not exactness)
    g->heavy_cap = (uint32_t)(n / kHeavy + 1);
    g->ht_size = 1024;
    while ((unsigned long long)g->ht_size < 4ull * g->heavy_cap) g->ht_size <<= 1;
    g->pool_cap = std::min<unsigned long long>(2ull * (unsigned long long)n + 65536ull, 0x7fffffffull);
    cudaError_t e;
    if ((e = cudaMalloc(&g->pts, (size_t)n * 3 * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc(&g->sorted, (size_t)n * sizeof(float4))) != cudaSuccess ||
        (e = cudaMalloc(&g->sorted_tmp, (size_t)n * sizeof(float4))) != cudaSuccess ||
        (e = cudaMalloc(&g->ht, (size_t)g->ht_size * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&g->pool, (size_t)g->pool_cap * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->heavy, (size_t)g->heavy_cap * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->tab, (g->ncell + 1) * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->code, (size_t)n * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->sums, (size_t)g->ntile * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&g->order_cnt, (size_t)((n + kOrderBlock - 1) / kOrderBlock) * sizeof(uint32_t))) != cudaSuccess ||
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
    cudaFree(g->sorted_tmp);
    cudaFree(g->ht);
    cudaFree(g->pool);
    cudaFree(g->heavy);
    cudaFree(g->tab);
    cudaFree(g->code);
    cudaFree(g->sums);
    cudaFree(g->order_cnt);
    cudaFree(g->bb);
    cudaFree(g->params);
    delete g;
    return DFB_OK;
}

extern "C" int dfb_grid_morton_order(const dfb_grid* g, int64_t q_begin, int64_t q_end, int32_t* d_out, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(g != nullptr, DFB_E_ARG, "dfb_grid_morton_order: grid is NULL");
    DFB_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= g->n, DFB_E_ARG,
                "dfb_grid_morton_order: range [%lld,%lld) outside the cloud", (long long)q_begin, (long long)q_end);
    if (q_end == q_begin) return DFB_OK;
    DFB_REQUIRE(d_out != nullptr, DFB_E_ARG, "dfb_grid_morton_order: d_out is NULL");
    cudaStream_t st = as_stream(stream);
    const long long nb = (g->n + kOrderBlock - 1) / kOrderBlock;
    order_count_kernel<<<(unsigned)nb, kOrderBlock, 0, st>>>(g->sorted, g->n, (int)q_begin, (int)q_end, g->order_cnt);
    scan_sums_serial<<<1, 1024, 0, st>>>(g->order_cnt, nb);
    order_scatter_kernel<<<(unsigned)nb, kOrderBlock, 0, st>>>(g->sorted, g->n, (int)q_begin, (int)q_end, g->order_cnt, d_out);
    note_launch(3);
    return check_cuda(cudaGetLastError(), "dfb_grid_morton_order kernels", __FILE__, __LINE__);
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
    a.ht = grid->ht; a.ht_mask = grid->ht_size - 1; a.pool = grid->pool;
    a.query = d_query; a.q_begin = q_begin; a.q_count = q_count;
    a.k = k; a.kcap = kcap;
    a.out_idx = d_idx; a.out_dist = d_dist; a.out_rad = d_rad;
    int k2 = 32;
    while (k2 < k) k2 <<= 1;
    // the candidate list (d2 and index, kcap entries per warp) + the bucket histogram of the final step (k2 ints per warp)
    const size_t smem = (size_t)kKnnWarps * (kcap * (sizeof(double) + sizeof(int)) + (k2 <= 512 && k2 / 32 >= DFB_KNN_BUCKET_MIN_E ? k2 : 0) * sizeof(int));
    // every warp works through a contiguous run of queries (at least 2): exactly one wave of resident CTAs, so that
    // all of them finish together
    long long blocks = (q_count + 2 * kKnnWarps - 1) / (2 * kKnnWarps);
#define DFB_KNN_LAUNCH(E)                                                                                         \
    do {                                                                                                          \
        if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&knn_kernel<E>), 100 * 1024)) return rc;       \
        int per_sm = 0;                                                                                           \
        DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_kernel<E>, kKnnWarps * 32, smem));    \
        const long long cap = (long long)sm_count() * (per_sm > 0 ? per_sm : 1);                                  \
        if (blocks > cap) blocks = cap;                                                                           \
        knn_kernel<E><<<(unsigned)blocks, kKnnWarps * 32, smem, as_stream(stream)>>>(a);                          \
    } while (0)
    switch (k2) {
        case 32: DFB_KNN_LAUNCH(1); break;
        case 64: DFB_KNN_LAUNCH(2); break;
        case 128: DFB_KNN_LAUNCH(4); break;
        case 256: DFB_KNN_LAUNCH(8); break;
        case 512: DFB_KNN_LAUNCH(16); break;
        default: DFB_KNN_LAUNCH(0); break;
    }
#undef DFB_KNN_LAUNCH
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
