// A C-ABI client of libdeepfit_b200.so with no Python and no torch in the loop: plain device pointers from the CUDA
// runtime, the entry points of include/deepfit_b200.h, results to stdout.  tests/test_gpu_c_client.py builds it with
// nvcc, runs it and compares with the Python binding on the same cloud.
//   usage: abi_client <cloud.f32 (n x 3 float32, raw)> <n> <k> <order> <n_query> [<weights.f32> <precision>]
// Without weights: the stage-by-stage calls (kNN -> patch/PCA -> classic jet fit).  With weights (the 20 folded layers
// of enum dfb_layer_id, each [out x in] weights then [out] bias, raw float32): dfb_model_create + dfb_pipeline_create +
// ONE dfb_pipeline_run in DeepFit mode -- the whole hot path from compiled code.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "deepfit_b200.h"

#define CK(x)                                                                                 \
    do {                                                                                      \
        int rc_ = (x);                                                                        \
        if (rc_ != 0) {                                                                       \
            std::fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, dfb_last_error_string());   \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)
#define CU(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));           \
            return 3;                                                               \
        }                                                                           \
    } while (0)

int main(int argc, char** argv) {
    if (argc != 6 && argc != 8) return 1;
    const long long n = std::atoll(argv[2]);
    const int k = std::atoi(argv[3]), order = std::atoi(argv[4]);
    const long long nq = std::atoll(argv[5]);
    std::vector<float> h_pts((size_t)n * 3);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(h_pts.data(), sizeof(float), h_pts.size(), f) != h_pts.size()) return 1;
    std::fclose(f);
    CK(dfb_device_check());
    float *d_pts, *d_patch, *d_U, *d_nrm;
    int32_t *d_idx, *d_info;
    double *d_rad, *d_curv;
    CU(cudaMalloc(&d_pts, h_pts.size() * sizeof(float)));
    CU(cudaMemcpy(d_pts, h_pts.data(), h_pts.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&d_idx, (size_t)nq * k * sizeof(int32_t)));
    CU(cudaMalloc(&d_rad, (size_t)nq * sizeof(double)));
    CU(cudaMalloc(&d_patch, (size_t)nq * 3 * k * sizeof(float)));
    CU(cudaMalloc(&d_U, (size_t)nq * 9 * sizeof(float)));
    CU(cudaMalloc(&d_nrm, (size_t)nq * 3 * sizeof(float)));
    CU(cudaMalloc(&d_curv, (size_t)nq * 2 * sizeof(double)));
    CU(cudaMalloc(&d_info, (size_t)nq * sizeof(int32_t)));
    dfb_grid* grid = nullptr;
    CK(dfb_grid_create(d_pts, n, nullptr, &grid));
    if (argc == 8) {
        static const int dims[DFB_NUM_LAYERS][2] = {{3, 64},    {64, 128},  {128, 1024}, {1024, 512}, {512, 256},  {256, 4},   {3, 64},
                                                    {64, 64},   {64, 64},   {64, 128},   {128, 1024}, {1024, 512}, {512, 256}, {256, 4096},
                                                    {64, 128},  {128, 1024}, {1088, 512}, {512, 256},  {256, 128},  {128, 1}};
        size_t total = 0;
        for (auto& d : dims) total += (size_t)d[0] * d[1] + d[1];
        std::vector<float> h_w(total);
        FILE* fw = std::fopen(argv[6], "rb");
        if (!fw || std::fread(h_w.data(), sizeof(float), total, fw) != total) return 1;
        std::fclose(fw);
        dfb_model_desc desc;
        desc.k = k;
        desc.weight_mode = DFB_WEIGHT_MODE_SIGMOID;
        desc.precision = std::atoi(argv[7]);
        desc.max_batch = 1024;
        size_t off = 0;
        for (int i = 0; i < DFB_NUM_LAYERS; ++i) {
            desc.layers[i].in_features = dims[i][0];
            desc.layers[i].out_features = dims[i][1];
            desc.layers[i].h_weight = h_w.data() + off;
            off += (size_t)dims[i][0] * dims[i][1];
            desc.layers[i].h_bias = h_w.data() + off;
            off += dims[i][1];
        }
        dfb_model* model = nullptr;
        CK(dfb_model_create(&desc, nullptr, &model));
        dfb_pipeline* pipe = nullptr;
        CK(dfb_pipeline_create(grid, model, k, order, 1024, 0, &pipe));
        dfb_pipeline_outputs o = {};
        o.d_normals = d_nrm;
        o.d_curv = order > 1 ? d_curv : nullptr;
        o.d_info = d_info;
        CK(dfb_pipeline_run(pipe, 0, nq, &o, nullptr));
        CK(dfb_model_status(model, nullptr));
        CU(cudaDeviceSynchronize());
        std::vector<float> h_nrm((size_t)nq * 3);
        std::vector<double> h_curv((size_t)nq * 2, 0.0);
        CU(cudaMemcpy(h_nrm.data(), d_nrm, h_nrm.size() * sizeof(float), cudaMemcpyDeviceToHost));
        if (order > 1) CU(cudaMemcpy(h_curv.data(), d_curv, h_curv.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < nq; ++i)
            std::printf("0 %.9g %.9g %.9g %.17g %.17g\n", h_nrm[i * 3], h_nrm[i * 3 + 1], h_nrm[i * 3 + 2], h_curv[i * 2], h_curv[i * 2 + 1]);
        CK(dfb_pipeline_destroy(pipe));
        CK(dfb_model_destroy(model));
        CK(dfb_grid_destroy(grid));
        std::fprintf(stderr, "abi %d, kernels launched %lld\n", dfb_abi_version(), (long long)dfb_launch_count(0));
        return 0;
    }
    CK(dfb_knn(grid, nullptr, 0, nq, k, d_idx, nullptr, d_rad, nullptr));
    CK(dfb_patch_pca(d_pts, n, d_idx, nullptr, 0, nq, k, d_rad, 1, d_patch, d_U, nullptr));
    CK(dfb_wls_fit(d_patch, nullptr, nullptr, d_U, d_rad, nq, k, order, nullptr, nullptr, d_nrm, nullptr,
                   order > 1 ? d_curv : nullptr, nullptr, nullptr, d_info, nullptr));
    CU(cudaDeviceSynchronize());
    std::vector<float> h_nrm((size_t)nq * 3);
    std::vector<double> h_curv((size_t)nq * 2, 0.0);
    std::vector<int32_t> h_idx((size_t)nq * k);
    CU(cudaMemcpy(h_nrm.data(), d_nrm, h_nrm.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (order > 1) CU(cudaMemcpy(h_curv.data(), d_curv, h_curv.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_idx.data(), d_idx, h_idx.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (long long i = 0; i < nq; ++i)
        std::printf("%d %.9g %.9g %.9g %.17g %.17g\n", h_idx[(size_t)i * k + k - 1], h_nrm[i * 3], h_nrm[i * 3 + 1], h_nrm[i * 3 + 2],
                    h_curv[i * 2], h_curv[i * 2 + 1]);
    CK(dfb_grid_destroy(grid));
    std::fprintf(stderr, "abi %d, kernels launched %lld\n", dfb_abi_version(), (long long)dfb_launch_count(0));
    return 0;
}
