/* rgc_features.h — C-ABI of the B200-native A-LOAM feature path (part of librgc_gicp.so).
 *
 * Replaces the numeric body of ScanRegistration::laserCloudHandler
 * (/root/reference/rgc_slam/src/scanRegistration.cpp:110-663 and removeClosedPointCloud :732-763):
 * range gate, ring id / relative time, ring-ordered cloud, range, incidence angle, near-range
 * intensity smoothing, the three curvatures, ground marking + weighted ground plane, occlusion
 * masking, and the per-(ring, sextant) sort + greedy edge / planar / intensity-edge selection.
 * ROS I/O (:104-109, :687-727) is out of scope.  Scans are processed in batches (one CUDA block
 * per scan / per ring), because one scan is far too small to occupy a B200.
 *
 * Layout: the batch is one concatenated array of raw points (x, y, z, intensity; 4 floats, firing
 * order) plus `scan_offsets[n_scans + 1]`.  Every per-point output array has the SAME capacity
 * layout with 8 slots of slack per scan: scan b owns [scan_offsets[b] + 8 b, scan_offsets[b+1] + 8 (b+1)),
 * of which the first cloud_size[b] entries are valid (the ring-ordered cloud is never longer than
 * the raw scan).  Feature lists are per-scan fixed-capacity blocks in the reference's push_back
 * order, holding indices into that scan's ordered cloud.  Any output pointer may be NULL.
 */
#ifndef RGC_FEATURES_H
#define RGC_FEATURES_H

#include <stddef.h>
#include <stdint.h>

#include "rgc_gicp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  const float* xyzi;           /* host, 4 floats per raw point */
  const int32_t* scan_offsets; /* host, n_scans + 1 entries (in points) */
  int n_scans;
  int n_rings;          /* N_SCANS: 16, 32 or 64        (scanRegistration.cpp:29,57) */
  double minimum_range; /* MINIMUM_RANGE (0.5)          (:31,:59) */
  double maximum_range; /* MAXMUM_RANGE (launch: 80)    (:31,:60) */
  int use_intensity;    /* USE_intensity (1)            (:30,:58) */
} rgc_scan_batch;

/* capacities of the per-scan feature-list blocks, in entries */
#define RGC_FEAT_CAP_SHARP(n_rings) ((n_rings)*6 * 20)      /* cornerPointsSharp  (:493-504) */
#define RGC_FEAT_CAP_LESS_SHARP(n_rings) ((n_rings)*6 * 22) /* cornerPointsLessSharp (:503,:508,:617) */
#define RGC_FEAT_CAP_FLAT(n_rings) ((n_rings)*6 * 40)       /* surfPointsFlat     (:546-556) */
#define RGC_FEAT_CAP_INTEN(n_rings) ((n_rings)*6 * 20)      /* intenPointsSharp   (:601-611) */
#define RGC_FEAT_CAP_LESS_INTEN(n_rings) ((n_rings)*6 * 21) /* intenPointsLessSharp (:611,:616) */

typedef struct {
  /* ---- per scan ---- */
  int32_t* cloud_size;  /* n_scans: points kept after range gate + ring assignment (:215) */
  int32_t* scan_start;  /* n_scans x 64: scanStartInd (:223) */
  int32_t* scan_end;    /* n_scans x 64: scanEndInd   (:229) */
  double* groundparam;  /* n_scans x 11: normal, vector1, vector2, distance, source (:420-430) */
  int32_t* ground_size; /* n_scans: number of (duplicated) ground samples (:346) */
  int32_t* inten_merged; /* n_scans: 1 if intenPointsSharp is appended to cornerPointsSharp (:653-656) */
  /* ---- per ordered point (capacity layout above) ---- */
  float* cloud;       /* 4 floats: x, y, z, scanID + 0.1 * relTime (:207-211) */
  int32_t* src_index; /* index of the point in its raw scan */
  int32_t* intensity_num; /* smoothed integer intensity (:257-268) */
  float* range_vec;       /* (:234-237) */
  float* scan_angle;      /* (:239-255) */
  float* curvature;       /* cloudCurvature (:279) */
  float* inten_curvature; /* intensityCurvature (:285,:290) */
  float* curvature2;      /* cloudCurvature2 (:294) */
  float* distance_source; /* (:280) */
  float* other_source;    /* (:292) */
  int32_t* label;         /* cloudLabel: 2 sharp, 1 less sharp, -1 flat, 0 other */
  int32_t* inten_label;   /* intenLabel */
  int32_t* neighbor_picked;
  int32_t* inten_neighbor_picked;
  int32_t* ground_marked;
  /* ---- per scan feature lists (indices into the scan's ordered cloud) + counts (n_scans) ---- */
  int32_t* corner_sharp;      float* corner_sharp_w; int32_t* n_corner_sharp;      /* normal_x = distance_source + 1 (:501) */
  int32_t* corner_less_sharp;                        int32_t* n_corner_less_sharp;
  int32_t* surf_flat;         float* surf_flat_w;    int32_t* n_surf_flat;         /* normal_x = distance_source (:554) */
  int32_t* inten_sharp;       float* inten_sharp_w;  int32_t* n_inten_sharp;       /* normal_x = other_source (:609) */
  int32_t* inten_less_sharp;                         int32_t* n_inten_less_sharp;
  /* ---- the two unbounded clouds of the handler, as index lists ---- */
  int32_t* surf_less_flat;   /* surfPointsLessFlatScan (:586-592): per-point capacity layout (like `label`), ascending */
  int32_t* n_surf_less_flat; /* n_scans */
  int32_t* ground_points;    /* GroundPoints (:338), the cloud published on /laser_cloud_ground (:714-717): n_scans x ground_cap
                                indices in push order, duplicates included; scan b holds min(ground_size[b], ground_cap) entries */
  int32_t ground_cap;        /* slots per scan in ground_points (a sample is appended up to 10 times: 10 x the points of rings 0-6 at most) */
  float device_ms; /* out: CUDA-event time of all kernels of this call */
} rgc_feat_out;

/* One pass of the feature path over a batch of scans.  Returns RGC_OK or an rgc_status error
 * (e.g. RGC_ERR_UNSUPPORTED for n_rings not in {16,32,64} or a ring segment longer than 2048). */
int rgc_feat_extract(rgc_ctx* ctx, const rgc_scan_batch* batch, rgc_feat_out* out);

#ifdef __cplusplus
}
#endif
#endif /* RGC_FEATURES_H */
