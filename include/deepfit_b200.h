/*
 * deepfit_b200 -- C ABI of the B200-native (sm_100a) DeepFit inference hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (sitzikbs/DeepFit, pure Python) has no FFI of its own; each entry point below replaces the
 * Python/PyTorch function cited next to it (paths relative to the reference root), and is what a
 * ctypes binding on the reference side would bind (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every function returns int: 0 = ok, negative = usage error (DFB_E_*), positive = CUDA
 *     runtime error code; dfb_last_error_string() returns the thread-local message.
 *   - every pointer named d_* is a DEVICE pointer owned by the caller (the host side allocates
 *     with torch); the library never frees caller memory.  h_* are host pointers.
 *   - every call takes the CUDA stream to launch on and is asynchronous w.r.t. the host unless
 *     stated otherwise.  Handles are immutable after creation except for their private workspace,
 *     so one handle must not be used from two streams at once.
 *   - there is no CPU fallback: without an sm_100 device the compute calls fail with DFB_E_NODEVICE.
 */
#ifndef DEEPFIT_B200_H_
#define DEEPFIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_ABI_VERSION 2

#if defined(__GNUC__)
#define DFB_API __attribute__((visibility("default")))
#else
#define DFB_API
#endif

#define DFB_OK 0
#define DFB_E_ARG (-1)          /* bad argument (null pointer, size out of range) */
#define DFB_E_UNSUPPORTED (-2)  /* valid in the reference but not supported here (see message) */
#define DFB_E_NODEVICE (-3)     /* no CUDA device of compute capability 10.x is current */
#define DFB_E_INTERNAL (-4)     /* a kernel reported an internal fault (e.g. pipeline time-out) */

typedef void* dfb_stream; /* cudaStream_t */
typedef struct dfb_grid dfb_grid;
typedef struct dfb_model dfb_model;

DFB_API int dfb_abi_version(void);
DFB_API const char* dfb_last_error_string(void);
/* 0 when the current device is compute capability 10.x; DFB_E_NODEVICE otherwise. */
DFB_API int dfb_device_check(void);

/* Number of CUDA kernels this library has launched in this process (reset != 0 zeroes the counter
 * after reading). */
DFB_API int64_t dfb_launch_count(int32_t reset);

/* ------------------------------------------------------------------------------------------
 * Stage 1 -- neighbourhood search.
 * Replaces scipy.spatial.cKDTree(points, 10) built at tutorial/tutorial_utils.py:16 (and
 * dataset.py:32) and its per-point query(points[i], k) at tutorial/tutorial_utils.py:21
 * (dataset.py:293).
 * ------------------------------------------------------------------------------------------ */

/* Build the search structure (Morton-ordered multi-level uniform grid) over d_points f32[n,3].
 * The points are copied (and re-ordered) into the handle; d_points may be released afterwards.
 * Allocates; the build kernels themselves are asynchronous on the stream. */
DFB_API int dfb_grid_create(const float* d_points, int64_t n, dfb_stream stream, dfb_grid** out);
/* Rebuild in place for a new cloud of the SAME size (no allocation, no host synchronisation). */
DFB_API int dfb_grid_update(dfb_grid* grid, const float* d_points, dfb_stream stream);
DFB_API int dfb_grid_destroy(dfb_grid* grid);
/* The point indices i with q_begin <= i < q_end in the grid's spatial (Morton) order: d_out int32[q_end - q_begin].
 * The reference visits queries in file order (tutorial/compute_normals.py:56, DataLoader without shuffle); queries are
 * independent, so a caller may visit them in any order -- in this one consecutive queries share their neighbourhood
 * cells, which is what dfb_knn's hints exploit (dfb_pipeline_run does so and scatters the results back to file order). */
DFB_API int dfb_grid_morton_order(const dfb_grid* grid, int64_t q_begin, int64_t q_end, int32_t* d_out, dfb_stream stream);

/* Exact k nearest neighbours of q_count query points.  Queries are points of the cloud:
 * d_query == NULL -> point indices q_begin .. q_begin+q_count-1, else d_query[0..q_count) (q_begin ignored).
 * Ranking key: (fp64 squared distance accumulated x->y->z on fp32 coordinates, point index),
 * i.e. cKDTree's result after distance-then-index tie-breaking.
 *   d_idx  i32[q_count,k]  neighbour indices, ascending by the key (the query itself first)
 *   d_dist f64[q_count,k]  Euclidean distances sqrt(d2), may be NULL
 *   d_rad  f64[q_count]    distance of the k-th neighbour (`rad = max(point_distances)`,
 *                          tutorial_utils.py:22), may be NULL
 * Requires 1 <= k <= min(n, 1024). */
DFB_API int dfb_knn(const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count, int32_t k,
            int32_t* d_idx, double* d_dist, double* d_rad, dfb_stream stream);

/* Fixed-radius neighbourhoods: every point with d2 <= r*r (d2 as in dfb_knn), i.e. the SET that
 * cKDTree.query_ball_point returns (reference dataset.py:289-291, neighbor_search_method='r'; the reference's
 * element ORDER is the tree's traversal order and is not reproduced).  Two passes:
 *   dfb_ball_count  d_count i64[q_count] = neighbours per query;
 *   (caller: exclusive scan -> d_offsets i64[q_count], allocates d_idx i32[sum])
 *   dfb_ball_fill   d_idx[d_offsets[q] ...] = the neighbour indices of query q, in a deterministic grid order.
 * Radius: d_radius f64[q_count] per query, or NULL -> the scalar `radius` for all. */
DFB_API int dfb_ball_count(const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count,
                   const double* d_radius, double radius, int64_t* d_count, dfb_stream stream);
DFB_API int dfb_ball_fill(const dfb_grid* grid, const int32_t* d_query, int64_t q_begin, int64_t q_count,
                  const double* d_radius, double radius, const int64_t* d_offsets, int32_t* d_idx,
                  dfb_stream stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2 -- patch centring, scaling and PCA alignment.
 * Replaces SinglePointCloudDataset.__getitem__ lines tutorial/tutorial_utils.py:23-30 and
 * pca_points :35-57 (dataset.py:320-366 for the PCPNet loader).
 *   patch = (points[idx] - points[centre]) / rad  (fp32), U = principal axes by descending
 *   variance (sign: largest-magnitude component of each axis positive), output patch * U.
 *   d_patch f32[q_count,3,k] (planar x-row, y-row, z-row: the reference's [B,3,k])
 *   d_U     f32[q_count,3,3] (`trans`)
 * flags: bit 0 = divide by rad (0: keep world scale, d_rad may be NULL); bit 1 = skip the PCA
 * (U = identity, patch = centred/scaled points), the reference's use_pca=False.
 * ------------------------------------------------------------------------------------------ */
DFB_API int dfb_patch_pca(const float* d_points, int64_t n_points, const int32_t* d_idx, const int32_t* d_query,
                  int64_t q_begin, int64_t q_count, int32_t k, const double* d_rad, int32_t flags,
                  float* d_patch, float* d_U, dfb_stream stream);
/* Same with ragged neighbourhoods (ball queries with fewer than k points, reference dataset.py:297-366): only the
 * first d_valid[q] entries of a row of d_idx are real neighbours; the remaining rows of the patch are zero padding
 * and take no part in the PCA mean / covariance.  d_valid == NULL -> every row is full. */
DFB_API int dfb_patch_pca_ragged(const float* d_points, int64_t n_points, const int32_t* d_idx, const int32_t* d_valid,
                         const int32_t* d_query, int64_t q_begin, int64_t q_count, int32_t k, const double* d_rad,
                         int32_t scale_by_rad, float* d_patch, float* d_U, dfb_stream stream);


/* ------------------------------------------------------------------------------------------
 * Stage 3 -- the point-weight network (QSTN + feature STN + PointNet encoder + weight head).
 * Replaces DeepFit.forward up to the weights, models/DeepFit.py:301-316 (PointNetFeatures
 * :180-208, PointNetEncoder :229-237, STN :353-380, QSTN :410-441, batch_quat_to_rotmat
 * utils/normal_estimation_utils.py:365-390), eval-mode BatchNorm folded into the preceding layer
 * by the host.  arch='simple', use_point_stn=1, use_feat_stn=True, point_tuple=1, sym_op='max'.
 * ------------------------------------------------------------------------------------------ */
enum dfb_layer_id {
    DFB_L_QSTN_CONV1 = 0, /* 3    -> 64   feat.pointfeat.stn1.conv1 + bn1 */
    DFB_L_QSTN_CONV2,     /* 64   -> 128  ...conv2 + bn2 */
    DFB_L_QSTN_CONV3,     /* 128  -> 1024 ...conv3 + bn3 */
    DFB_L_QSTN_FC1,       /* 1024 -> 512  ...fc1 + bn4 */
    DFB_L_QSTN_FC2,       /* 512  -> 256  ...fc2 + bn5 */
    DFB_L_QSTN_FC3,       /* 256  -> 4    ...fc3 */
    DFB_L_PF_CONV1,       /* 3    -> 64   feat.pointfeat.conv1 + bn1 */
    DFB_L_PF_CONV2,       /* 64   -> 64   feat.pointfeat.conv2 + bn2 */
    DFB_L_STN_CONV1,      /* 64   -> 64   feat.pointfeat.stn2.conv1 + bn1 */
    DFB_L_STN_CONV2,      /* 64   -> 128 */
    DFB_L_STN_CONV3,      /* 128  -> 1024 */
    DFB_L_STN_FC1,        /* 1024 -> 512 */
    DFB_L_STN_FC2,        /* 512  -> 256 */
    DFB_L_STN_FC3,        /* 256  -> 4096 */
    DFB_L_ENC_CONV2,      /* 64   -> 128  feat.conv2 + bn2 */
    DFB_L_ENC_CONV3,      /* 128  -> 1024 feat.conv3 + bn3 (no ReLU) */
    DFB_L_HEAD_CONV1,     /* 1088 -> 512  conv1 + bn1 (columns 0..1023 global feature, 1024..1087 point feature) */
    DFB_L_HEAD_CONV2,     /* 512  -> 256  conv2 + bn2 */
    DFB_L_HEAD_CONV3,     /* 256  -> 128  conv3 + bn3 */
    DFB_L_HEAD_CONV4,     /* 128  -> 1    conv4 */
    DFB_NUM_LAYERS
};

typedef struct dfb_layer {
    const float* h_weight; /* HOST fp32 [out_features, in_features] row-major, BatchNorm already folded */
    const float* h_bias;   /* HOST fp32 [out_features] */
    int32_t in_features;
    int32_t out_features;
} dfb_layer;

#define DFB_WEIGHT_MODE_SIGMOID 0 /* w = 0.01 + sigmoid(a)      models/DeepFit.py:316 */
#define DFB_WEIGHT_MODE_TANH 1    /* w = (1.01 + tanh(a)) / 2   models/DeepFit.py:313-314 */

/* Operand precision of the tensor-core products (all accumulate in fp32).  Weight error vs the reference's fp32 network on
 * the golden patches / what that does to BASELINE.json's gates (0.01 deg, 1e-3 curvature) -- tests/test_gpu_network.py: */
#define DFB_PRECISION_BF16X3 0 /* split BF16, 3 products, 16-bit operands: |dw| 1e-4; passes both gates on smooth clouds,
                                  up to 1.2x over the curvature gate on ill-conditioned lidar patches */
#define DFB_PRECISION_BF16 1   /* single BF16 product: 3x fewer tensor ops, ~0.2 deg -- an approximation, not parity */
#define DFB_PRECISION_F16F8 2  /* FP16 main product + ONE E4M3 MMA carrying both first-order corrections (2 product
                                  units): |dw| 3e-4; passes on smooth clouds, up to 2.9x over the curvature gate on lidar
                                  patches; needs |weights| < 250 and |activations| < 500 (checked) */
#define DFB_PRECISION_F16X3 3  /* split FP16, 3 products, 22-bit operands: |dw| 2e-5, which is the tensor core's truncating
                                  fp32 accumulate, not the operands (the reference's own fp32 network is 2e-6 from the
                                  fp64 one); THE parity mode (default): every per-row gate holds;
                                  same range requirements as F16F8 (checked) */

#define DFB_PRECISION_F16MIX 4 /* F16X3 operands everywhere, but the chain layers named by a DFB_MIX_* mask issue only the
                                  hi*hi product (11-bit operands, a third of the tensor work): the per-layer precision
                                  ladder.  precision = DFB_PRECISION_F16MIX | (mask << 8); mask 0 = DFB_MIX_DEFAULT, which
                                  is 1.21x faster than F16X3 with |dw| 4e-4: passes both gates on smooth clouds (0.17 /
                                  0.53 of the gates), up to 7x over the curvature gate on ill-conditioned lidar patches,
                                  where no single-product layer passes (measured: profiles/r2_mix_probe.log) */
#define DFB_MIX_QSTN_SMALL 1   /* QSTN chain 64 -> 128 */
#define DFB_MIX_QSTN_MAX 2     /* QSTN chain 128 -> 1024 + max-pool */
#define DFB_MIX_STN_SMALL 4    /* feature-STN chain 64 -> 64 -> 128 (and 3 -> 64 -> 64 recomputed in front of it) */
#define DFB_MIX_STN_MAX 8      /* feature-STN chain 128 -> 1024 + max-pool */
#define DFB_MIX_ENC_SMALL 16   /* encoder chain 64 -> 64, x * T2, 64 -> 128 */
#define DFB_MIX_ENC_MAX 32     /* encoder chain 128 -> 1024 + max-pool */
#define DFB_MIX_ALL 63
#define DFB_MIX_DEFAULT (DFB_MIX_QSTN_MAX | DFB_MIX_STN_MAX)

typedef struct dfb_model_desc {
    int32_t k;           /* points per patch == num_points of the reference constructor; 128, 256, 384 or 512 */
    int32_t weight_mode; /* DFB_WEIGHT_MODE_* */
    int32_t precision;   /* DFB_PRECISION_* */
    int32_t max_batch;   /* patches per internal chunk (workspace is sized for it); 0 = default */
    dfb_layer layers[DFB_NUM_LAYERS];
} dfb_model_desc;

/* Packs the weights into the tensor-core operand layout on the device and allocates the
 * activation workspace.  Synchronous. */
DFB_API int dfb_model_create(const dfb_model_desc* desc, dfb_stream stream, dfb_model** out);
DFB_API int dfb_model_destroy(dfb_model* model);

/* d_patch f32[n,3,k] -> d_weights f32[n,k], d_trans f32[n,3,3] (QSTN rotation `trans`),
 * d_trans2 f32[n,64,64] (feature transform, may be NULL), d_points_rot f32[n,3,k] (patch * trans,
 * the points the jet is fitted to, may be NULL). */
DFB_API int dfb_model_forward(dfb_model* model, const float* d_patch, int64_t n, float* d_weights, float* d_trans,
                      float* d_trans2, float* d_points_rot, dfb_stream stream);

/* Synchronises the stream and reports whether any tensor-core pipeline of this model timed out
 * (DFB_E_INTERNAL); kernels never spin forever on a barrier. */
DFB_API int dfb_model_status(dfb_model* model, dfb_stream stream);

/* Per-kernel-class device timing of dfb_model_forward (CUDA events on the launching stream).
 * enable != 0 starts recording; dfb_model_profile_read synchronises, adds the elapsed milliseconds and
 * launch counts of each class to ms[]/launches[] (n_classes entries, see DFB_PROF_*) and clears the
 * records. */
#define DFB_PROF_POINTWISE 0  /* K=3 layers + rotation (CUDA cores) */
#define DFB_PROF_GEMM_SMALL 1 /* 64->64, 64->128 and x*T2 (tensor cores) */
#define DFB_PROF_GEMM_MAX 2   /* 128->1024 + max-pool (tensor cores, dominant) */
#define DFB_PROF_FC 3         /* per-patch fully connected layers */
#define DFB_PROF_HEAD 4       /* 64->512, 512->256, 256->128 + weight epilogue */
#define DFB_PROF_QUAT 5       /* fc3 + quaternion -> rotation */
#define DFB_PROF_NUM 6
DFB_API int dfb_model_profile(dfb_model* model, int32_t enable);
DFB_API int dfb_model_profile_read(dfb_model* model, double* ms, int64_t* launches, int32_t n_classes, dfb_stream stream);

/* ------------------------------------------------------------------------------------------
 * Stage 4+5 -- weighted n-jet fit, normal, principal curvatures.
 * Replaces fit_Wjet + solve_linear_system (models/DeepFit.py:11-154), the normal at :88,
 * compute_principal_curvatures (tutorial/tutorial_utils.py:196-226 == models/DeepFit.py:444-474)
 * and the caller epilogue tutorial/compute_normals.py:67,79 / test_n_est.py:154-161.
 *   d_patch   f32[n,3,k]  PCA-aligned patch
 *   d_weights f32[n,k]    NULL = classic unweighted fit (compute_normals.py:72)
 *   d_trans   f32[n,3,3]  NULL or the QSTN rotation: the jet is fitted to patch * trans and the
 *                         world normal is rotated back by trans^T
 *   d_U       f32[n,3,3]  NULL or the PCA frame: n_world = (n_local * trans^T) * U^T
 *   d_rad     f64[n]      NULL or patch radius: curv_scaled = curv / rad (fp64, compute_normals.py:79)
 * outputs (each may be NULL):
 *   d_beta f32[n,M] (M = 3,6,10,15 for order 1..4), d_n_local f32[n,3], d_n_world f32[n,3],
 *   d_curv f32[n,2] ascending (order >= 2 only), d_curv_scaled f64[n,2],
 *   d_dirs f32[n,2,3]: the second return value of compute_principal_curvatures (tutorial_utils.py:223-224):
 *                  torch.symeig's eigenvector matrix (column j belongs to curvature j) with a zero third
 *                  column; eigenvector signs are library-defined in the reference, here the largest-magnitude
 *                  component of each eigenvector is positive,
 *   d_dirs_world f32[n,2,3]: its rows rotated back by trans^T and U^T as test_c_est.py:162-168 does,
 *   d_info i32[n]: 0 ok, 1 = Cholesky failed and the deterministic ridge (diag *= 1.01) was used,
 *                  2 = ridge failed too (beta = 0).
 * All n rows are solved (the reference leaves rows >= 16*floor(n/16) at zero, :129-133).
 * ------------------------------------------------------------------------------------------ */
DFB_API int dfb_wls_fit(const float* d_patch, const float* d_weights, const float* d_trans, const float* d_U,
                const double* d_rad, int64_t n, int32_t k, int32_t order, float* d_beta, float* d_n_local,
                float* d_n_world, float* d_curv, double* d_curv_scaled, float* d_dirs, float* d_dirs_world,
                int32_t* d_info, dfb_stream stream);

/* Principal curvatures and directions from jet coefficients alone (order >= 2): d_beta f32[n,M] ->
 * d_curv f32[n,2] (may be NULL), d_dirs f32[n,2,3] (may be NULL; layout and sign rule as in dfb_wls_fit).
 * Replaces compute_principal_curvatures (tutorial/tutorial_utils.py:196-226). */
DFB_API int dfb_principal_curvatures(const float* d_beta, int64_t n, int32_t m, float* d_curv, float* d_dirs,
                                     dfb_stream stream);

/* Analytic normals of the fitted jet at every neighbour: fit_Wjet(..., compute_neighbor_normals=True),
 * models/DeepFit.py:90-115 (gradient evaluated at the preconditioned x/h, y/h with the un-preconditioned
 * beta, as the reference does).  d_patch f32[n,3,k], d_trans NULL or f32[n,3,3] (the jet was fitted to
 * patch * trans), d_beta f32[n,M] -> d_out f32[n,k,3]. */
DFB_API int dfb_neighbor_normals(const float* d_patch, const float* d_trans, const float* d_beta, int64_t n, int32_t k,
                                 int32_t order, float* d_out, dfb_stream stream);

/* ------------------------------------------------------------------------------------------
 * The fused path 1 -> 5 for a range of point indices of the cloud a grid was built over.
 * Replaces the batch loop of tutorial/compute_normals.py:56-81 (DataLoader over SinglePointCloudDataset,
 * model forward, undoing the rotations :67 / test_n_est.py:154-161, curv / rad :79, torch.cat) and, with the
 * optional outputs, the per-batch body of test_c_est.py:125-168.
 *   model NULL -> classic unweighted fit (compute_normals.py:72); else DeepFit mode with the model's k.
 *   chunk     queries per internal chunk (0 = default: 16 384 DeepFit, 262 144 classic); the workspace
 *             (neighbour indices, patches, frames; double-buffered) is sized for it at creation.
 *   flags     DFB_PIPE_KEEP_QSTN: do not undo the QSTN rotation on the world normal / directions, which is
 *             what the shipped tutorial script does (compute_normals.py:67 applies only the PCA frame).
 * dfb_pipeline_run is asynchronous on `stream`; internally stages 1+2 of the next chunk run on a second
 * stream owned by the pipeline.  One pipeline must not run on two streams at once.  The grid and the model
 * must outlive the pipeline.  Outputs are indexed from q_begin; every pointer but d_normals may be NULL.
 * ------------------------------------------------------------------------------------------ */
typedef struct dfb_pipeline dfb_pipeline;
#define DFB_PIPE_KEEP_QSTN 1
#define DFB_PIPE_FILE_ORDER 2 /* visit the queries in file order (the reference's) instead of the grid's Morton order;
                                 same results, slower neighbour search */
typedef struct dfb_pipeline_outputs {
    float* d_normals;    /* f32[q,3]  world normals (unoriented), compute_normals.py:67 */
    double* d_curv;      /* f64[q,2]  principal curvatures / rad, ascending, compute_normals.py:79 (order >= 2) */
    float* d_beta;       /* f32[q,M]  jet coefficients */
    float* d_weights;    /* f32[q,k]  point weights (DeepFit mode) */
    float* d_trans;      /* f32[q,9]  QSTN rotation (DeepFit mode) */
    float* d_trans2;     /* f32[q,64,64] feature transform (DeepFit mode) */
    float* d_U;          /* f32[q,9]  PCA frame of the patch (the dataset's `trans`) */
    double* d_rad;       /* f64[q]    patch radius */
    float* d_dirs_world; /* f32[q,2,3] principal directions rotated to world space, test_c_est.py:162-168 */
    int32_t* d_info;     /* i32[q]    solver status as in dfb_wls_fit */
} dfb_pipeline_outputs;
DFB_API int dfb_pipeline_create(const dfb_grid* grid, dfb_model* model, int32_t k, int32_t order, int32_t chunk,
                                int32_t flags, dfb_pipeline** out);
DFB_API int dfb_pipeline_run(dfb_pipeline* pipeline, int64_t q_begin, int64_t q_end, const dfb_pipeline_outputs* outputs,
                             dfb_stream stream);
DFB_API int dfb_pipeline_destroy(dfb_pipeline* pipeline);

/* ------------------------------------------------------------------------------------------
 * Host-side text I/O (all host threads; no GPU needed).  Byte-compatible with numpy's defaults.
 * dfb_savetxt_*: replaces np.savetxt(path, array) (tutorial/compute_normals.py:87,91;
 *   test_c_est.py:201-210): "%.18e", ' ' delimiter, '\n' newline.  float32 values are widened first,
 *   as numpy does.
 * dfb_loadtxt_f32: replaces np.loadtxt(path).astype('float32') (tutorial/tutorial_utils.py:13):
 *   call with h_out == NULL to get rows/cols, then with a buffer of rows*cols floats.
 * ------------------------------------------------------------------------------------------ */
DFB_API int dfb_savetxt_f32(const char* path, const float* h_data, int64_t rows, int32_t cols);
DFB_API int dfb_savetxt_f64(const char* path, const double* h_data, int64_t rows, int32_t cols);
DFB_API int dfb_loadtxt_f32(const char* path, float* h_out, int64_t capacity, int64_t* rows_out, int32_t* cols_out);

#ifdef __cplusplus
}
#endif
#endif /* DEEPFIT_B200_H_ */
