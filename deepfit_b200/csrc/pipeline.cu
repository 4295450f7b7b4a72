// The whole hot path for a range of point indices behind ONE C call: kNN -> centre / scale / PCA -> weight network ->
// weighted jet fit -> normal + curvature (+ principal directions).
//
// Replaces the batch loop of the reference's tutorial/compute_normals.py:56-81 (DataLoader over
// SinglePointCloudDataset, model forward, back-rotation, curv / rad, torch.cat) and, with the optional outputs, the
// per-batch body of test_c_est.py:125-168.  A C / C++ host binding the ABI does not have to re-write the chunk loop, the
// workspace management or the stream order.
//
// Schedule: queries are processed in chunks.  Stages 1+2 (kNN, PCA) of chunk i+1 run on a second, lowest-priority
// stream owned by the pipeline while stages 3-5 (network, jet fit) of chunk i run on the caller's stream; the
// neighbourhood buffers are double-buffered and handed over with events, so the search fills the SMs the persistent
// tensor-core kernels leave idle in their tails instead of serialising in front of them.
//
// Classic mode (no network) splits stage 2: the gather kernel writes the centred / scaled patch UNROTATED plus its
// covariance, a one-thread-per-patch kernel turns the covariance into the PCA frame U (the serial fp64 Jacobi sweeps ran
// 32 times redundantly per patch inside the fused warp-per-patch kernel), and the jet-fit kernel applies U while it reads
// the patch, through the per-patch rotation input it has for the QSTN -- the same fmaf chain as the fused kernel's, so
// the results are bit-identical (DFB_PIPE_FUSED_PCA=1 restores the fused kernel for A/B runs).  In DeepFit mode the
// network needs the rotated patch, so the fused kernel stays.  Both gather through the grid's float4 copy of the cloud.
//
// Visiting order: queries are independent, so the range is walked in the grid's spatial (Morton) order
// (dfb_grid_morton_order) -- consecutive queries then share their neighbourhood cells, which dfb_knn's level and radius
// hints turn into 1.5x the search rate -- and every result is written to its FILE-order row (the jet-fit kernel scatters
// its own outputs; the pass-through outputs go through a row-scatter kernel).  DFB_PIPE_FILE_ORDER restores the
// reference's visiting order (tutorial/compute_normals.py:56, DataLoader without shuffle); results are bit-identical.
#include "common.cuh"

#include <stdlib.h>

namespace {
// dst[(out_row[i] - base) * row + j] = src[i * row + j]
template <typename T>
__global__ void scatter_rows_kernel(const T* __restrict__ src, long long rows, int row, const int* __restrict__ out_row, long long base,
                                    T* __restrict__ dst) {
    const long long n = rows * row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / row;
        dst[((long long)out_row[r] - base) * row + (i - r * row)] = src[i];
    }
}
template <typename T>
int scatter_rows(const T* src, long long rows, int row, const int* out_row, long long base, T* dst, cudaStream_t st) {
    const long long n = rows * row;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    scatter_rows_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(src, rows, row, out_row, base, dst);
    dfb::note_launch();
    return dfb::check_cuda(cudaGetLastError(), "scatter_rows_kernel launch", __FILE__, __LINE__);
}
}  // namespace

struct dfb_pipeline {
    const dfb_grid* grid = nullptr;
    dfb_model* model = nullptr;   // NULL = classic (unweighted) fit
    int k = 0, order = 0, chunk = 0, flags = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_start = nullptr, ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    // per-chunk workspace, [2] = double-buffered between the two streams
    int32_t* idx[2] = {nullptr, nullptr};
    double* rad[2] = {nullptr, nullptr};
    float* patch[2] = {nullptr, nullptr};
    float* U[2] = {nullptr, nullptr};
    double* cov[2] = {nullptr, nullptr};   // classic mode: [chunk,6] patch covariances (split stage 2)
    bool split_pca = false, gather4 = true;
    float* weights = nullptr;     // [chunk,k]   (network -> jet fit, caller's stream only)
    float* trans = nullptr;       // [chunk,9]
    float* trans2 = nullptr;      // [chunk,4096] staging, only when trans2 is requested in Morton order
    int32_t* qorder = nullptr;    // the query range in Morton order (capacity qorder_cap)
    long long qorder_cap = 0;
    int device = 0;
};

extern "C" int dfb_pipeline_destroy(dfb_pipeline* p) {
    if (!p) return DFB_OK;
    for (int b = 0; b < 2; ++b) {
        cudaFree(p->idx[b]);
        cudaFree(p->rad[b]);
        cudaFree(p->patch[b]);
        cudaFree(p->U[b]);
        cudaFree(p->cov[b]);
        if (p->ev_ready[b]) cudaEventDestroy(p->ev_ready[b]);
        if (p->ev_free[b]) cudaEventDestroy(p->ev_free[b]);
    }
    cudaFree(p->weights);
    cudaFree(p->trans);
    cudaFree(p->trans2);
    cudaFree(p->qorder);
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    if (p->side) cudaStreamDestroy(p->side);
    delete p;
    return DFB_OK;
}

extern "C" int dfb_pipeline_create(const dfb_grid* grid, dfb_model* model, int32_t k, int32_t order, int32_t chunk,
                                   int32_t flags, dfb_pipeline** out) {
    using namespace dfb;
    DFB_REQUIRE(out != nullptr, DFB_E_ARG, "dfb_pipeline_create: out is NULL");
    *out = nullptr;
    DFB_REQUIRE(grid != nullptr, DFB_E_ARG, "dfb_pipeline_create: grid is NULL");
    DFB_REQUIRE(order >= 1 && order <= 4, DFB_E_ARG, "Polynomial order unsupported, please use 1..4 (got %d)", order);
    DFB_REQUIRE(k >= 1 && k <= 1024 && (long long)k <= grid_size(grid), DFB_E_ARG, "dfb_pipeline_create: k=%d out of range", k);
    DFB_REQUIRE(model == nullptr || model_k(model) == k, DFB_E_ARG,
                "dfb_pipeline_create: the network was built for %d points per patch (MaxPool1d(num_points), reference "
                "models/DeepFit.py:336,395), got k=%d", model ? model_k(model) : 0, k);
    DFB_REQUIRE(chunk >= 0, DFB_E_ARG, "dfb_pipeline_create: negative chunk");
    int rc = dfb_device_check();
    if (rc) return rc;
    dfb_pipeline* p = new dfb_pipeline();
    p->grid = grid;
    p->model = model;
    p->k = k;
    p->order = order;
    p->chunk = chunk > 0 ? chunk : (model ? 16384 : 262144);
    p->flags = flags;
    {
        const char* e1 = getenv("DFB_PIPE_FUSED_PCA");
        p->split_pca = model == nullptr && !(e1 && e1[0] == '1');
        const char* e2 = getenv("DFB_PIPE_GATHER3");      // A/B: gather through the f32[n,3] array (three loads per point)
        p->gather4 = !(e2 && e2[0] == '1');
    }
    cudaError_t e = cudaGetDevice(&p->device);
    int lo = 0, hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least priority
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, lo);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming);
    const size_t c = (size_t)p->chunk;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaEventCreateWithFlags(&p->ev_ready[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_free[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&p->idx[b], c * k * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc(&p->rad[b], c * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&p->patch[b], c * 3 * k * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->U[b], c * 9 * sizeof(float));
        if (e == cudaSuccess && p->split_pca) e = cudaMalloc(&p->cov[b], c * 6 * sizeof(double));
    }
    if (e == cudaSuccess && model) {
        e = cudaMalloc(&p->weights, c * k * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&p->trans, c * 9 * sizeof(float));
    }
    if (e != cudaSuccess) {
        dfb_pipeline_destroy(p);
        return check_cuda(e, "dfb_pipeline_create allocations", __FILE__, __LINE__);
    }
    *out = p;
    return DFB_OK;
}

extern "C" int dfb_pipeline_run(dfb_pipeline* p, int64_t q_begin, int64_t q_end, const dfb_pipeline_outputs* o, dfb_stream stream) {
    using namespace dfb;
    DFB_REQUIRE(p != nullptr && o != nullptr, DFB_E_ARG, "dfb_pipeline_run: NULL argument");
    DFB_REQUIRE(q_begin >= 0 && q_end >= q_begin && q_end <= grid_size(p->grid), DFB_E_ARG,
                "dfb_pipeline_run: query range [%lld,%lld) outside the cloud", (long long)q_begin, (long long)q_end);
    DFB_REQUIRE(o->d_normals != nullptr || q_end == q_begin, DFB_E_ARG, "dfb_pipeline_run: d_normals is NULL");
    DFB_REQUIRE(p->order > 1 || (o->d_curv == nullptr && o->d_dirs_world == nullptr), DFB_E_ARG,
                "Can't compute curvatures for jet with order less than 2");
    DFB_REQUIRE(p->model != nullptr || (o->d_weights == nullptr && o->d_trans == nullptr), DFB_E_ARG,
                "dfb_pipeline_run: weights / trans outputs need a network (classic mode has neither)");
    if (q_end == q_begin) return DFB_OK;
    cudaStream_t main_st = as_stream(stream);
    const int k = p->k;
    static const int kM[5] = {0, 3, 6, 10, 15};
    const int M = kM[p->order];
    const float* pts = grid_points(p->grid);
    const float4* pts4 = p->gather4 ? grid_points4(p->grid) : nullptr;
    const long long n_pts = grid_size(p->grid);
    const long long n_chunks = (q_end - q_begin + p->chunk - 1) / p->chunk;

    // the visiting order (on the caller's stream, before the side stream is released)
    const bool morton = !(p->flags & DFB_PIPE_FILE_ORDER);
    if (morton) {
        if (p->qorder_cap < q_end - q_begin) {
            DFB_CUDA(cudaFree(p->qorder));   // synchronises: no earlier run can still be reading it
            p->qorder = nullptr;
            p->qorder_cap = 0;
            DFB_CUDA(cudaMalloc(&p->qorder, (size_t)(q_end - q_begin) * sizeof(int32_t)));
            p->qorder_cap = q_end - q_begin;
        }
        if (o->d_trans2 && !p->trans2) DFB_CUDA(cudaMalloc(&p->trans2, (size_t)p->chunk * 4096 * sizeof(float)));
        int rc0 = dfb_grid_morton_order(p->grid, q_begin, q_end, p->qorder, stream);
        if (rc0) return rc0;
    }
    // the side stream starts after everything already queued on the caller's stream (e.g. a grid update)
    DFB_CUDA(cudaEventRecord(p->ev_start, main_st));
    DFB_CUDA(cudaStreamWaitEvent(p->side, p->ev_start, 0));

    auto front = [&](long long ci) -> int {   // stages 1+2 of chunk ci, on the side stream
        const int b = (int)(ci & 1);
        const long long s = q_begin + ci * p->chunk;
        const long long c = (q_end - s < p->chunk) ? (q_end - s) : p->chunk;
        if (ci >= 2) DFB_CUDA(cudaStreamWaitEvent(p->side, p->ev_free[b], 0));   // chunk ci-2 has been consumed
        const int32_t* ql = morton ? p->qorder + (s - q_begin) : nullptr;
        int rc = dfb_knn(p->grid, ql, s, c, k, p->idx[b], nullptr, p->rad[b], p->side);
        if (rc) return rc;
        rc = patch_pca_impl(pts, pts4, n_pts, p->idx[b], nullptr, ql, s, c, k, p->rad[b], 1, p->patch[b], p->U[b],
                            p->split_pca ? p->cov[b] : nullptr, p->side);
        if (rc) return rc;
        DFB_CUDA(cudaEventRecord(p->ev_ready[b], p->side));
        return 0;
    };

    int rc = front(0);
    if (rc) return rc;
    for (long long ci = 0; ci < n_chunks; ++ci) {
        const int b = (int)(ci & 1);
        const long long s = q_begin + ci * p->chunk;
        const long long c = (q_end - s < p->chunk) ? (q_end - s) : p->chunk;
        const long long off = s - q_begin;
        if (ci + 1 < n_chunks) {
            rc = front(ci + 1);
            if (rc) return rc;
        }
        DFB_CUDA(cudaStreamWaitEvent(main_st, p->ev_ready[b], 0));
        float* w = nullptr;
        float* tr = nullptr;
        const int32_t* rows = morton ? p->qorder + off : nullptr;   // file-order row of every patch of the chunk
        if (p->model) {
            // in Morton order the network writes chunk-local staging rows, scattered below when they are outputs
            w = (o->d_weights && !morton) ? o->d_weights + off * k : p->weights;
            tr = (o->d_trans && !morton) ? o->d_trans + off * 9 : p->trans;
            float* t2 = o->d_trans2 ? (morton ? p->trans2 : o->d_trans2 + off * 4096) : nullptr;
            rc = dfb_model_forward(p->model, p->patch[b], c, w, tr, t2, nullptr, main_st);
            if (rc) return rc;
            if (morton) {
                if (o->d_weights && (rc = scatter_rows(w, c, k, rows, q_begin, o->d_weights, main_st))) return rc;
                if (o->d_trans && (rc = scatter_rows(tr, c, 9, rows, q_begin, o->d_trans, main_st))) return rc;
                if (o->d_trans2 && (rc = scatter_rows(t2, c, 4096, rows, q_begin, o->d_trans2, main_st))) return rc;
            }
        }
        const long long oo = morton ? 0 : off;   // with a row list the jet fit takes the base pointers
        // split stage 2 (classic): the patch is unrotated, U enters as the rotation applied on the fly and undone on the outputs
        rc = wls_fit_impl(p->patch[b], w, p->split_pca ? p->U[b] : tr, p->split_pca ? nullptr : p->U[b], p->rad[b], c, k, p->order, o->d_beta ? o->d_beta + oo * M : nullptr, nullptr,
                          o->d_normals + oo * 3, nullptr, (p->order > 1 && o->d_curv) ? o->d_curv + oo * 2 : nullptr, nullptr,
                          o->d_dirs_world ? o->d_dirs_world + oo * 6 : nullptr, o->d_info ? o->d_info + oo : nullptr,
                          (!p->split_pca && (p->flags & DFB_PIPE_KEEP_QSTN)) ? 1 : 0, main_st, rows, q_begin);
        if (rc) return rc;
        if (morton) {
            if (o->d_rad && (rc = scatter_rows(p->rad[b], c, 1, rows, q_begin, o->d_rad, main_st))) return rc;
            if (o->d_U && (rc = scatter_rows(p->U[b], c, 9, rows, q_begin, o->d_U, main_st))) return rc;
        } else {
            if (o->d_rad) DFB_CUDA(cudaMemcpyAsync(o->d_rad + off, p->rad[b], (size_t)c * sizeof(double), cudaMemcpyDeviceToDevice, main_st));
            if (o->d_U) DFB_CUDA(cudaMemcpyAsync(o->d_U + off * 9, p->U[b], (size_t)c * 9 * sizeof(float), cudaMemcpyDeviceToDevice, main_st));
        }
        DFB_CUDA(cudaEventRecord(p->ev_free[b], main_st));
    }
    return DFB_OK;
}
