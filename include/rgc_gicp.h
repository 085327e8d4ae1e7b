/* rgc_gicp.h — C-ABI of the B200-native scan-matching path (librgc_gicp.so).
 *
 * Drop-in boundary for the reference's CPU FastGICP path.  Each entry point names the
 * reference interface it replaces (paths under /root/reference/rgc_slam/):
 *   FG  = include/fast_gicp/gicp/          FGI = include/fast_gicp/gicp/impl/
 *   SRC = src/
 * The C++ facade include/rgc/fast_gicp.hpp wraps these 1:1 behind the pcl::Registration-style
 * names the odometry node calls (SRC/RGC_odometer.cpp:998-1011).
 *
 * Conventions
 *   - every function returns 0 on success, <0 (rgc_status) on failure; rgc_last_error(ctx)
 *     gives the message.  No exceptions, no torch types, caller-owned host memory in and out.
 *   - 4x4 / 6x6 matrices are COLUMN-major (Eigen's default, so Eigen::Matrix4f::data() passes
 *     straight through).  Covariances are arrays of column-major 4x4 doubles (Eigen::Matrix4d),
 *     row/column 3 zero, exactly the reference's std::vector<Eigen::Matrix4d>.
 *   - point clouds: pointer to the first point, count, byte stride; xyz are the first three
 *     floats of every point (pcl::PointXYZ 16 B, PointXYZI 32 B, PointNormal 48 B —
 *     SRC/fast_gicp/gicp/fast_gicp.cpp:4-6).  The homogeneous coordinate is taken as 1.
 *   - there is NO CPU fallback: with no usable CUDA device rgc_ctx_create fails.
 *   - threading: a context (its two CUDA streams, device-memory pool and pinned result buffers) and the
 *     objects created from it belong to ONE host thread at a time, like the reference's registration
 *     objects.  A context may serve any number of rgc_reg / rgc_map objects sequentially.  To run
 *     registrations concurrently on one GPU (batched loop-closure verification) give every host thread
 *     its own context: their kernels overlap on the device (tools/bench_c4.py).
 *   - asynchrony: rgc_reg_set_source / set_target return as soon as the upload and the first kernels of the
 *     cloud's build are ISSUED; the rest of the build (sort, voxel hash), the source k-NN / covariances and the
 *     host waits in between are driven by the first call that needs the cloud (align, linearize, get_*_covs,
 *     rgc_reg_sync_inputs ...), which interleaves source and target so each host wait overlaps the other
 *     cloud's device work.  Consequences: (1) a HOST buffer passed to a setter must stay valid and unchanged
 *     until that next call returns — exactly the lifetime the reference's shared_ptr gives the cloud
 *     (pageable memory is staged by the driver before the setter returns, pinned memory is read by DMA later);
 *     (2) a failure of the deferred part (non-finite coordinates, out of memory) is reported by that next call.
 *     Environment RGC_SYNC_BUILD=1 makes the setters complete the build before returning.
 *   - scratch blocks of a call go back to the context's pool on every exit path, failures included.
 *   - clouds must be finite: a NaN / inf coordinate makes set_source / set_target fail with
 *     RGC_ERR_INVALID (a NaN candidate would silently break the exactness of the k-NN).
 */
#ifndef RGC_GICP_H
#define RGC_GICP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rgc_ctx rgc_ctx;
typedef struct rgc_reg rgc_reg;

typedef enum {
  RGC_OK = 0,
  RGC_ERR_CUDA = -1,        /* a CUDA call failed (message has the CUDA error string) */
  RGC_ERR_INVALID = -2,     /* bad argument */
  RGC_ERR_STATE = -3,       /* e.g. align() before both clouds are set */
  RGC_ERR_UNSUPPORTED = -4, /* e.g. k_correspondences > 128 */
  RGC_ERR_NOMEM = -5
} rgc_status;

/* FG/gicp_settings.hpp:6 */
typedef enum { RGC_REG_NONE = 0, RGC_REG_MIN_EIG = 1, RGC_REG_NORMALIZED_MIN_EIG = 2, RGC_REG_PLANE = 3, RGC_REG_FROBENIUS = 4 } rgc_regularization;
/* FG/lsq_registration.hpp:13 */
typedef enum { RGC_OPT_GAUSS_NEWTON = 0, RGC_OPT_LEVENBERG_MARQUARDT = 1 } rgc_optimizer;

/* Defaults = FGI/lsq_registration_impl.hpp:9-22 and FGI/fast_gicp_impl.hpp:8-23. */
typedef struct {
  int max_iterations;                /* pcl setMaximumIterations            (64)      */
  double rotation_epsilon;           /* setRotationEpsilon                  (2e-3)    */
  double transformation_epsilon;     /* pcl setTransformationEpsilon        (5e-4)    */
  float max_correspondence_distance; /* pcl setMaxCorrespondenceDistance    (FLT_MAX) */
  int k_correspondences;             /* setCorrespondenceRandomness         (20), <= 128 (fast paths: <= 32) */
  int regularization;                /* setRegularizationMethod             (PLANE)   */
  int optimizer;                     /* lsq_optimizer_type_                 (LM)      */
  int lm_max_iterations;             /* lm_max_iterations_                  (10)      */
  double lm_init_lambda_factor;      /* setInitialLambdaFactor              (1e-9)    */
  int lm_debug_print;                /* setDebugPrint                       (0)       */
  float grid_cell;                   /* finest voxel edge of the kNN grid in metres; 0 = default (0.05) */
} rgc_params;

typedef struct {
  int converged;            /* pcl hasConverged()                                   */
  int iterations;           /* nr_iterations_ (FGI/lsq_registration_impl.hpp:66)    */
  int n_linearize;          /* device linearize launches                            */
  int n_compute_error;      /* device compute_error launches                        */
  int n_inliers;            /* correspondences within max distance at the last linearize */
  double final_error;       /* sum e^T M e at the last accepted linearize           */
  double final_hessian[36]; /* getFinalHessian(), column-major                      */
  float device_ms;          /* CUDA-event time of the whole align on the ctx stream */
} rgc_result;

/* ---- context: device, stream, pooled device memory, pinned result buffers ------------------ */
int rgc_ctx_create(int device, rgc_ctx** out);
int rgc_ctx_destroy(rgc_ctx* ctx);
const char* rgc_last_error(const rgc_ctx* ctx);
int rgc_ctx_synchronize(rgc_ctx* ctx);
/* raw cudaStream_t of the context (so callers can record CUDA events on the launching stream) */
void* rgc_ctx_stream(rgc_ctx* ctx);
/* profiling: when on, CUDA events bracket the kernels of every linearize / compute_error and
 * rgc_ctx_last_kernel_ms returns the device time of the last [k_correspond, k_linearize,
 * k_compute_error] launch in ms (live roofline numbers for bench.py).  Off by default. */
int rgc_ctx_set_profiling(rgc_ctx* ctx, int on);
int rgc_ctx_last_kernel_ms(const rgc_ctx* ctx, float* ms3);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t rgc_ctx_launch_count(const rgc_ctx* ctx);

/* ---- registration object: fast_gicp::FastGICP<PointSource, PointTarget> (FG/fast_gicp.hpp:20) -- */
int rgc_reg_create(rgc_ctx* ctx, rgc_reg** out); /* FastGICP() — cheap, tolerates per-frame create/destroy */
int rgc_reg_destroy(rgc_reg* reg);
void rgc_params_default(rgc_params* p);
int rgc_reg_set_params(rgc_reg* reg, const rgc_params* p);
int rgc_reg_get_params(const rgc_reg* reg, rgc_params* p);

/* setInputSource / setInputTarget (FGI/fast_gicp_impl.hpp:72-91).  `identity_key` plays the role
 * of the shared_ptr identity: non-zero and equal to the current key => early return, nothing is
 * copied or recomputed; otherwise the cloud is uploaded, its search structure rebuilt and its
 * covariances cleared (recomputed lazily inside align, FGI/fast_gicp_impl.hpp:104-109).  */
int rgc_reg_set_source(rgc_reg* reg, const void* points, size_t n, size_t stride_bytes, uint64_t identity_key);
int rgc_reg_set_target(rgc_reg* reg, const void* points, size_t n, size_t stride_bytes, uint64_t identity_key);
/* same, but `points` is a DEVICE pointer (cloud already resident in HBM) */
int rgc_reg_set_source_device(rgc_reg* reg, const void* d_points, size_t n, size_t stride_bytes, uint64_t identity_key);
int rgc_reg_set_target_device(rgc_reg* reg, const void* d_points, size_t n, size_t stride_bytes, uint64_t identity_key);

/* complete the deferred part of set_source / set_target now (see "asynchrony" above) and report its status */
int rgc_reg_sync_inputs(rgc_reg* reg);

int rgc_reg_swap_source_and_target(rgc_reg* reg); /* FGI/fast_gicp_impl.hpp:49-57 */
int rgc_reg_clear_source(rgc_reg* reg);           /* :59-63 */
int rgc_reg_clear_target(rgc_reg* reg);           /* :65-69 */

/* set/get{Source,Target}Covariances (FG/fast_gicp.hpp:62-72): n column-major 4x4 doubles */
int rgc_reg_set_source_covs(rgc_reg* reg, const double* m4x4, size_t n);
int rgc_reg_set_target_covs(rgc_reg* reg, const double* m4x4, size_t n);
int rgc_reg_get_source_covs(rgc_reg* reg, double* m4x4, size_t n); /* computes them if missing */
int rgc_reg_get_target_covs(rgc_reg* reg, double* m4x4, size_t n);

/* pcl::Registration::align(output, guess) -> FastGICP::computeTransformation
 * (FGI/fast_gicp_impl.hpp:103-112) -> LsqRegistration::computeTransformation
 * (FGI/lsq_registration_impl.hpp:53-79).  guess may be NULL (identity).  out_points (nullable):
 * n_source x 4 floats, the source transformed by the result (the `output` cloud).        */
int rgc_reg_align(rgc_reg* reg, const float* guess16, float* final_T16, rgc_result* result, float* out_points);

/* LsqRegistration::evaluateCost / linearize (FGI/lsq_registration_impl.hpp:48-51,
 * FGI/fast_gicp_impl.hpp:155-211).  H (36, column-major) and b (6) may both be NULL.        */
int rgc_reg_linearize(rgc_reg* reg, const double* T16, double* err, double* H36, double* b6);
/* FastGICP::compute_error (FGI/fast_gicp_impl.hpp:214-237): correspondences and Mahalanobis
 * matrices stay frozen from the last linearize.                                            */
int rgc_reg_compute_error(rgc_reg* reg, const double* T16, double* err);
/* correspondences_ / sq_distances_ of the last linearize, caller's index space (-1 = none;
 * sq_dist is +inf where no target lies within max_correspondence_distance)                  */
int rgc_reg_get_correspondences(rgc_reg* reg, int32_t* corr, float* sq_dist);

/* How the TARGET covariances of the exact-1-NN FastGICP path are produced.
 *   on_demand = 1 (default): FastGICP::linearize reads target_covs_[target_index] only at the current
 *     correspondences (FGI/fast_gicp_impl.hpp:139-146), so the k-NN + covariance of a target point is
 *     computed the first time it becomes a correspondence, by the same kernels with the parameters in
 *     force at the first align — the values linearize sees are bit-identical to the eager pass, but a
 *     500k-point submap costs <= n_source k-NN queries per align instead of 500k per frame.
 *   on_demand = 0: reference schedule, all n_target covariances at the first align (:107-109).
 * get_target_covs, the voxelised mode, swap_source_and_target and user-supplied covariances always see /
 * produce the complete set.  Environment RGC_EAGER_TARGET_COV=1 makes 0 the default.            */
int rgc_reg_set_target_covariance_mode(rgc_reg* reg, int on_demand);
/* profiling (rgc_ctx_set_profiling): device time of the on-demand k-NN + covariance kernels of the last linearize */
int rgc_ctx_last_ondemand_ms(const rgc_ctx* ctx, float* ms);

/* number of correspondences the last linearize used (points within max_correspondence_distance,
 * or (point, voxel) pairs in voxelised mode) */
int rgc_reg_last_inliers(const rgc_reg* reg, int* n);

/* pcl::Registration::getFitnessScore(max_range) with the final transformation of the last
 * align (callers: SRC/RGC_odometer.cpp:1010, SRC/RGC_mapping.cpp:2070)                        */
int rgc_reg_fitness(rgc_reg* reg, double max_range, double* score);
int rgc_reg_get_final_transformation(const rgc_reg* reg, float* T16);

/* pcl::search::KdTree::nearestKSearch for a batch of queries (exact; ascending by (d2, index)).
 * idx/d2: m x k row-major; rows are padded with -1 / +inf when k > n.  Test hook + public op. */
int rgc_knn(rgc_ctx* ctx, const void* points, size_t n, size_t stride_bytes, const void* queries, size_t m, size_t qstride_bytes, int k,
            int32_t* idx, float* d2, float grid_cell);

/* ---- voxelised GICP: fast_gicp::FastVGICP (FG/fast_vgicp.hpp:24, FGI/fast_vgicp_impl.hpp:17-204) — the
 * class SRC/RGC_odometer.cpp:998 instantiates.  When enabled, linearize / compute_error / align use
 * voxel correspondences (GaussianVoxelMap, FG/fast_vgicp_voxel.hpp:129-160) instead of the exact 1-NN:
 *   resolution        setResolution               (1.0)
 *   neighbor_search   setNeighborSearchMethod     0 = DIRECT27, 1 = DIRECT7, 2 = DIRECT1 (default)   (FG/gicp_settings.hpp:8)
 *   accumulation_mode setVoxelAccumulationMode    0 = ADDITIVE (default), 1 = ADDITIVE_WEIGHTED, 2 = MULTIPLICATIVE (:10)
 * max_correspondence_distance is ignored in this mode, as in the reference.                          */
int rgc_reg_set_vgicp(rgc_reg* reg, int enabled, double resolution, int neighbor_search, int accumulation_mode);
/* test hook: the Gaussian voxels of the target (unordered): coords 3 ints, num_points, mean 3 doubles,
 * covariance upper triangle 6 doubles per voxel; n_voxels is always set.                             */
int rgc_reg_get_voxels(rgc_reg* reg, int32_t* coords3, int32_t* num_points, double* mean3, double* cov6, size_t cap, size_t* n_voxels);

/* k nearest neighbours of every point of a cloud within the cloud itself (self included, rank 0),
 * through the same warp-cooperative kernel calculate_covariances uses (FGI/fast_gicp_impl.hpp:254).
 * idx: n x k row-major, original indices, ascending by (d2, index); -1 padding when k > n.      */
int rgc_knn_self(rgc_ctx* ctx, const void* points, size_t n, size_t stride_bytes, int k, int32_t* idx, float grid_cell);

/* ---- sharded target: this rank holds only a spatial slab of the target
 * (plus a halo >= max_correspondence_distance).  A source point is handled by exactly one rank: the
 * one whose slab [lo, hi) along `axis` (0/1/2) contains its TRANSFORMED position (computed in float,
 * identically on every rank).  axis < 0 switches sharding off.                                      */
int rgc_reg_set_owner_slab(rgc_reg* reg, int axis, float lo, float hi);
/* the same selection done by the library: `points` is the FULL target (host); the points with
 * lo - halo <= p[axis] < hi + halo are kept ON THE DEVICE (input order) and become this rank's target, and [lo, hi)
 * (+-inf allowed at the ends) becomes its ownership slab.  halo >= max_correspondence_distance + the k-NN radius of
 * the covariance neighbourhoods.  n_local / local_index (nullable: original index of each kept point) are outputs. */
int rgc_reg_set_target_slab(rgc_reg* reg, const void* points, size_t n, size_t stride_bytes, int axis, float lo, float hi, float halo,
                            uint64_t identity_key, size_t* n_local, int32_t* local_index);

/* Cross-rank sum of the partial (err, H, b).  Preferred: a communicator owned by the library — the partial sums
 * of linearize (29 doubles), compute_error (1), the two together when the LM loop issues them back to back (30) and
 * fitness (2) are summed by ONE launch on the context's stream and picked up by the host's poll: no host code runs
 * between the reduction kernels and the totals, and every rank takes identical LM steps.  Transport: each rank's
 * 1-block kernel stores its partials into a mailbox on EVERY rank (device memory mapped through CUDA IPC, plain
 * stores over NVLink, the call's sequence number packed into every word), polls its own mailbox for the other
 * ranks' words, adds the partials in rank order (bit-identical totals everywhere) and writes them to the host result area (rgc_comm_transport() == 1).  If
 * the mailboxes cannot be mapped on every rank (no peer access, world > 16, RGC_NO_P2P=1) the same sums are one
 * ncclAllReduce + a publishing kernel (rgc_comm_transport() == 0).  rank 0 calls rgc_comm_unique_id and hands the 128 bytes to the
 * other ranks by any side channel (torch.distributed, MPI, a file); every rank then calls rgc_comm_create
 * (collective).  libnccl.so.2 is opened at run time (RGC_NCCL_LIB overrides the name).                       */
typedef struct rgc_comm rgc_comm;
int rgc_comm_unique_id(char* id128);
int rgc_comm_create(rgc_ctx* ctx, const char* id128, int rank, int world, rgc_comm** out);
int rgc_comm_destroy(rgc_comm* comm);
int rgc_comm_info(const rgc_comm* comm, int* rank, int* world, uint64_t* n_allreduce, int* nccl_version);
int rgc_comm_transport(const rgc_comm* comm); /* 1 = peer-memory mailboxes (k_peer_allreduce), 0 = ncclAllReduce */
int rgc_reg_set_comm(rgc_reg* reg, rgc_comm* comm); /* NULL switches it off */
/* mean latency (us, CUDA events) of `reps` back-to-back all-reduces of n doubles on the context's stream; collective */
int rgc_comm_allreduce_us(rgc_comm* comm, int n_doubles, int reps, float* us);

/* Alternative: a caller-supplied reduction.  After every linearize / compute_error / fitness kernel the partial
 * sums (n doubles: 29 / 1 / 2) sit in `d_buf` (device memory, caller-owned, >= 29 doubles); `fn` must sum them
 * across ranks IN PLACE on the context's stream before returning 0 (gloo in the CPU tests, any other transport).
 * fn == NULL switches the hook off.                                                                            */
typedef int (*rgc_reduce_fn)(void* user, void* d_buf, int n_doubles);
int rgc_reg_set_allreduce(rgc_reg* reg, rgc_reduce_fn fn, void* user, void* d_buf);

/* per-stage device times (ms, CUDA events) of the last set_source / set_target / align on this reg:
 * [0] source build (ingest+sort+tables) [1] source kNN [2] source cov [3] target build
 * [4] target kNN [5] target cov [6] LM loop (all linearize/compute_error incl. host turnarounds) */
int rgc_reg_stage_ms(const rgc_reg* reg, float* ms7);

#ifdef __cplusplus
}
#endif
#endif /* RGC_GICP_H */
