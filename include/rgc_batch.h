/* rgc_batch.h — batched independent registrations (BASELINE.json configs[3]): loop-closure
 * candidate verification.  The reference verifies ONE candidate per 1 Hz tick
 * (SRC/RGC_mapping.cpp:1962-2086: detectLoopClosure -> pcl::IterativeClosestPoint::align -> accept iff
 * hasConverged() && getFitnessScore() <= historyKeyframeFitnessScore, :2070-2071).  This entry point runs B
 * such registrations — each a fast_gicp::FastGICP align of its own (source, target) pair from its own guess —
 * as ONE job: all clouds of a chunk are sorted into two multi-cloud voxel hashes (sources / targets), the k-NN
 * + covariances of all sources are one launch, and every LM round is one set of launches over all pairs still
 * iterating (correspondences, on-demand target covariances, linearize, compute_error), with per-pair partial
 * sums reduced in the same fixed order as the single-registration kernels.  The 6x6 solves and the LM step
 * control of the B pairs run on the host between rounds, pair by pair, exactly as FGI/lsq_registration_impl.hpp
 * :53-172 prescribes; pairs that have converged drop out of the launches.
 *
 * Results are bit-identical to calling rgc_reg_set_target / rgc_reg_set_source / rgc_reg_align /
 * rgc_reg_fitness on each pair separately (tests/test_gpu_batch.py).  No CPU fallback.
 */
#ifndef RGC_BATCH_H
#define RGC_BATCH_H

#include "rgc_gicp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one registration of the batch: host pointers, PCL-style strides (xyz = first three floats of a point) */
typedef struct {
  const void* source;
  size_t n_source, source_stride;
  const void* target;
  size_t n_target, target_stride;
  float guess[16]; /* column-major 4x4, the `guess` of pcl::Registration::align */
} rgc_pair;

typedef struct {
  float final_T[16]; /* getFinalTransformation(), column-major */
  rgc_result result; /* as rgc_reg_align fills it (device_ms = the chunk's device time / pairs in the chunk) */
  double fitness;    /* getFitnessScore(fitness_max_range); only if want_fitness */
} rgc_pair_result;

/* Align `n_pairs` independent pairs with the parameters `prm` (NULL = defaults, FGI/lsq_registration_impl.hpp
 * :9-22, FGI/fast_gicp_impl.hpp:8-23).  Pairs are processed in chunks of at most `max_chunk_pairs` (0 = choose:
 * as many as keep the chunk under ~24 M points).  Input buffers must stay valid until the call returns. */
int rgc_batch_align(rgc_ctx* ctx, const rgc_params* prm, const rgc_pair* pairs, size_t n_pairs, int want_fitness, double fitness_max_range,
                    int max_chunk_pairs, rgc_pair_result* out);

/* device time (ms, CUDA events) of the stages of the last rgc_batch_align on this context, summed over its
 * chunks: [0] upload + ingest [1] source build [2] target build [3] source kNN + covariances [4] LM rounds
 * (correspondences, on-demand target kNN + covariances, linearize, compute_error, host turnarounds) [5] fitness;
 * rounds = number of LM rounds launched */
int rgc_batch_last_stage_ms(const rgc_ctx* ctx, float* ms6, int* rounds);

#ifdef __cplusplus
}
#endif
#endif /* RGC_BATCH_H */
