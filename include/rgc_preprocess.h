/* rgc_preprocess.h — C-ABI of the step in front of the registration (part of librgc_gicp.so).
 *
 * Replaces, on the GPU and without the clouds leaving HBM between the stages:
 *   - RGC_odometer::adjustDistortion for the full cloud
 *     (/root/reference/rgc_slam/src/RGC_odometer.cpp:1441-1481, third loop): per-point motion
 *     compensation with s = 1 - frac(intensity) / SCAN_PERIOD (float arithmetic, :323),
 *     q = Identity.slerp(s, q_last_curr^-1), p' = q * (p - s * t_last_curr);
 *   - the two pcl::VoxelGrid<PointType> centroid filters in front of the FastVGICP call
 *     (RGC_odometer.cpp:975-991, leaf 0.2 m for the source and 0.3 m for the submap; defaults
 *     downsample_all_data = true, min_points_per_voxel = 0; output in ascending voxel-index order);
 *   - and, fused, setInputSource / setInputTarget on the filtered cloud (:1007-1008).
 *
 * Points are records of `stride` bytes with x, y, z (float) at byte 0 and the intensity (float) at
 * byte `intensity_offset` (16 for pcl::PointXYZI; RGC_NO_INTENSITY if the type has none, in which case
 * the filter averages zeros and de-skew is refused).  Outputs are packed (x, y, z, intensity) floats.
 * Points must be finite (the reference removes NaNs in scanRegistration before these stages; PCL's
 * is_dense = false filtering of non-finite points is not reproduced).
 * Same conventions as rgc_gicp.h: int status, rgc_last_error(ctx), host memory in and out.
 */
#ifndef RGC_PREPROCESS_H
#define RGC_PREPROCESS_H

#include <stddef.h>
#include <stdint.h>

#include "rgc_gicp.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RGC_NO_INTENSITY ((size_t)-1)

/* pcl::VoxelGrid<PointXYZI>::filter.  out_xyzi holds `cap` points (may be NULL to only count);
 * *n_out = number of occupied voxels (may exceed cap: call again with a larger buffer).
 * *passthrough (may be NULL) is set when the leaf is so small that PCL's int32 voxel index would
 * overflow: PCL then warns and returns the input cloud unchanged, and so does this. */
int rgc_voxel_grid(rgc_ctx* ctx, const void* points, size_t n, size_t stride, size_t intensity_offset, float leaf, float* out_xyzi, size_t cap,
                   size_t* n_out, int* passthrough);

/* adjustDistortion for one cloud.  q_wxyz = q_last_curr (w, x, y, z), t3 = t_last_curr, scan_period =
 * SCAN_PERIOD (0.1).  out_xyzi: n x 4 floats, same order as the input. */
int rgc_deskew(rgc_ctx* ctx, const void* points, size_t n, size_t stride, size_t intensity_offset, const double* q_wxyz, const double* t3,
               float scan_period, float* out_xyzi);

/* [de-skew if q_wxyz != NULL] -> [voxel filter if leaf > 0] -> setInputSource / setInputTarget, all on the
 * device.  identity_key as in rgc_reg_set_source (same key: nothing is recomputed).  *n_out (may be NULL)
 * = size of the cloud the registration now holds. */
int rgc_reg_set_source_filtered(rgc_reg* reg, const void* points, size_t n, size_t stride, size_t intensity_offset, float leaf, const double* q_wxyz,
                                const double* t3, float scan_period, uint64_t identity_key, size_t* n_out);
int rgc_reg_set_target_filtered(rgc_reg* reg, const void* points, size_t n, size_t stride, size_t intensity_offset, float leaf, const double* q_wxyz,
                                const double* t3, float scan_period, uint64_t identity_key, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* RGC_PREPROCESS_H */
