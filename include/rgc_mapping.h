/* rgc_mapping.h — C-ABI of the mapping node's scan-to-map association (part of librgc_gicp.so).
 *
 * Replaces, per LiDAR frame and Ceres iteration, the four association loops of
 * /root/reference/rgc_slam/src/RGC_mapping.cpp (:1093-1136 and :1139-1189 edge features of the current
 * and of the last frame, :1192-1240 and :1243-1290 planar features): pointAssociateToMap (:1811-1820),
 * pcl::KdTreeFLANN::nearestKSearch(k = 5) in the corner / surface map (:1097, :1196), the line test on
 * the scatter matrix of the 5 neighbours and the plane fit by colPivHouseholderQr.  What it returns are
 * the arguments of LidarEdgeFactor::Create(curr_point, point_a, point_b, weight) and
 * LidarPlaneNormFactor::Create(curr_point, norm, negative_OA_dot_norm, weight); the Ceres problem itself
 * stays on the host (north_star).  A map object plays kdtree*FromMap->setInputCloud (:1073-1074) and is
 * reused for all loops and both solver iterations of a frame.
 *
 * Feature points are records of `stride` bytes with x, y, z (float) at byte 0 (pcl::PointXYZINormal: 48).
 * Outputs are dense per-feature arrays in input order; entry i is meaningful iff valid[i] != 0 — the
 * residual blocks are added for the valid features in ascending i, as the reference does.
 */
#ifndef RGC_MAPPING_H
#define RGC_MAPPING_H

#include <stddef.h>
#include <stdint.h>

#include "rgc_gicp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rgc_map rgc_map;

int rgc_map_create(rgc_ctx* ctx, const void* map_points, size_t n, size_t stride, rgc_map** out);
int rgc_map_destroy(rgc_map* map);

/* q_wxyz, t3: q_w_curr (w, x, y, z) and t_w_curr (or the last frame's pose for the "last" loops).
 * point_a, point_b: n x 3 doubles.  *n_valid (may be NULL) = number of valid features (corner_num). */
int rgc_map_associate_edges(rgc_map* corner_map, const void* features, size_t n, size_t stride, const double* q_wxyz, const double* t3, int32_t* valid,
                            double* point_a, double* point_b, size_t* n_valid);
/* norm: n x 3 doubles (unit plane normal), negative_OA_dot_norm: n doubles. */
int rgc_map_associate_planes(rgc_map* surf_map, const void* features, size_t n, size_t stride, const double* q_wxyz, const double* t3, int32_t* valid,
                             double* norm, double* negative_OA_dot_norm, size_t* n_valid);

#ifdef __cplusplus
}
#endif
#endif /* RGC_MAPPING_H */
