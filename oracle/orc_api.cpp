// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.hpp header).  PARITY UNPINNED.
// Flat C entry points so tests / bench.py's cpu_baseline leg can drive the CPU restatement
// through ctypes.  Nothing in the product library links this file.
#include <chrono>
#include <cstring>

#include "orc_features.hpp"
#include "orc_gicp.hpp"
#include "orc_mapping.hpp"
#include "orc_preprocess.hpp"

using namespace orc;

template <class T>
static void cp(T* dst, const std::vector<T>& v) {
  if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(T) * v.size());
}

extern "C" {

// ---- mapping-node association ----
void orc_assoc_edges(const float* map, int nm, const float* feats, int n, const double* q, const double* t, int* valid, double* pa, double* pb) {
  MapAssoc m;
  m.build(map, nm);
  m.edges(feats, n, q, t, valid, pa, pb);
}
void orc_assoc_planes(const float* map, int nm, const float* feats, int n, const double* q, const double* t, int* valid, double* norm, double* dist) {
  MapAssoc m;
  m.build(map, nm);
  m.planes(feats, n, q, t, valid, norm, dist);
}
void orc_colpiv_qr_solve_5x3(const double* A, const double* b, double* x) { colpiv_qr_solve_5x3(A, b, x); }

// ---- pre-step: de-skew and pcl::VoxelGrid ----
void orc_deskew(const float* xyzi, int n, const double* q_wxyz, const double* t3, float scan_period, float* out) { deskew(xyzi, n, q_wxyz, t3, scan_period, out); }
int orc_voxel_grid(const float* xyzi, int n, float leaf, float* out, int* passthrough) { return voxel_grid(xyzi, n, leaf, out, passthrough); }

// ---- linear algebra probes (checked against numpy/scipy in tests/test_oracle_linalg.py) ----
void orc_jacobi_svd3(const double* A, double* U, double* sv, double* V) { jacobi_svd3(A, U, sv, V); }
void orc_eigh3(const double* A, double* evals, double* V) { eigh3(A, evals, V); }
void orc_inverse4(const double* A, double* inv) { inverse4(A, inv); }
void orc_ldlt6_solve(const double* A, const double* rhs, double* x) { ldlt6_solve(A, rhs, x); }
void orc_so3_exp(const double* omega, double* R9) {
  double q[4];
  so3_exp_quat(omega, q);
  quat_to_rot(q, R9);
}

// ---- kNN ----
// pts: n x 4 floats; queries: m x 4 floats; out idx/d2: m x k (rows padded with -1 / inf if k > n)
void orc_knn(const float* pts, int n, const float* queries, int m, int k, int* idx, float* d2, int num_threads) {
  KdTree tree;
  tree.build(pts, n);
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = -1;
      d2[(size_t)i * k + j] = INFINITY;
    }
    tree.knn(&queries[4 * (size_t)i], k, &idx[(size_t)i * k], &d2[(size_t)i * k]);
  }
}

// O(n*m) brute force with the same float arithmetic and (d2, idx) order; cross-checks the tree.
void orc_knn_bruteforce(const float* pts, int n, const float* queries, int m, int k, int* idx, float* d2) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = -1;
      d2[(size_t)i * k + j] = INFINITY;
    }
    KdTree::Result res{std::min(k, n), 0, &d2[(size_t)i * k], &idx[(size_t)i * k]};
    const float* q = &queries[4 * (size_t)i];
    for (int p = 0; p < n; p++) {
      float dx = q[0] - pts[4 * (size_t)p], dy = q[1] - pts[4 * (size_t)p + 1], dz = q[2] - pts[4 * (size_t)p + 2];
      float d = dx * dx;
      d = d + dy * dy;
      d = d + dz * dz;
      res.insert(d, p);
    }
  }
}

// covariances from given neighbour lists (k indices per point) — lets tests isolate A2 from A1
void orc_covariances_from_knn(const float* pts, int n, const int* knn_idx, int k, int method, double* covs16) {
#pragma omp parallel for schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    int found = 0;
    while (found < k && knn_idx[(size_t)i * k + found] >= 0) found++;
    FastGICP::covariance_from_neighbors(pts, &knn_idx[(size_t)i * k], found, k, method, &covs16[16 * (size_t)i]);
  }
}

// ---- FastGICP object ----
void* orc_gicp_create() { return new FastGICP(); }
void orc_gicp_destroy(void* h) { delete (FastGICP*)h; }
void orc_gicp_set_params(void* h, int max_iterations, double rotation_epsilon, double transformation_epsilon, float corr_dist, int k,
                         int regularization, int optimizer, int lm_max_iterations, double lm_init_lambda_factor, int num_threads) {
  FastGICP* g = (FastGICP*)h;
  g->max_iterations_ = max_iterations;
  g->rotation_epsilon_ = rotation_epsilon;
  g->transformation_epsilon_ = transformation_epsilon;
  g->corr_dist_threshold_ = corr_dist;
  g->k_correspondences_ = k;
  g->regularization_method_ = regularization;
  g->lsq_optimizer_type_ = optimizer;
  g->lm_max_iterations_ = lm_max_iterations;
  g->lm_init_lambda_factor_ = lm_init_lambda_factor;
  g->num_threads_ = num_threads > 0 ? num_threads : omp_get_max_threads();
}
void orc_gicp_set_source(void* h, const float* xyzw, int n) { ((FastGICP*)h)->setInputSource(xyzw, n); }
void orc_gicp_set_target(void* h, const float* xyzw, int n) { ((FastGICP*)h)->setInputTarget(xyzw, n); }
void orc_gicp_ensure_covariances(void* h) { ((FastGICP*)h)->ensure_covariances(); }
void orc_gicp_set_source_covs(void* h, const double* c16, int n) {
  FastGICP* g = (FastGICP*)h;
  g->source_covs_.resize(n);
  std::memcpy(g->source_covs_.data(), c16, sizeof(double) * 16 * (size_t)n);
}
void orc_gicp_set_target_covs(void* h, const double* c16, int n) {
  FastGICP* g = (FastGICP*)h;
  g->target_covs_.resize(n);
  std::memcpy(g->target_covs_.data(), c16, sizeof(double) * 16 * (size_t)n);
}
int orc_gicp_get_source_covs(void* h, double* c16) {
  FastGICP* g = (FastGICP*)h;
  std::memcpy(c16, g->source_covs_.data(), sizeof(double) * 16 * g->source_covs_.size());
  return (int)g->source_covs_.size();
}
int orc_gicp_get_target_covs(void* h, double* c16) {
  FastGICP* g = (FastGICP*)h;
  std::memcpy(c16, g->target_covs_.data(), sizeof(double) * 16 * g->target_covs_.size());
  return (int)g->target_covs_.size();
}
// trans row-major double 4x4; H row-major 6x6 (nullable with b)
double orc_gicp_linearize(void* h, const double* trans, double* H, double* b) {
  FastGICP* g = (FastGICP*)h;
  g->ensure_covariances();
  return g->linearize(trans, H, b);
}
double orc_gicp_compute_error(void* h, const double* trans) { return ((FastGICP*)h)->compute_error(trans); }
void orc_gicp_get_correspondences(void* h, int* corr, float* sq_dist) {
  FastGICP* g = (FastGICP*)h;
  std::memcpy(corr, g->correspondences_.data(), sizeof(int) * g->correspondences_.size());
  if (sq_dist) std::memcpy(sq_dist, g->sq_distances_.data(), sizeof(float) * g->sq_distances_.size());
}
// returns wall seconds spent inside align(); result: [converged, nr_iterations, n_linearize, n_compute_error]
double orc_gicp_align(void* h, const float* guess, float* final_T, float* out_points, int* result4, double* final_hessian) {
  FastGICP* g = (FastGICP*)h;
  g->n_linearize_ = g->n_compute_error_ = 0;
  auto t0 = std::chrono::steady_clock::now();
  g->align(guess, out_points);
  auto t1 = std::chrono::steady_clock::now();
  std::memcpy(final_T, g->final_transformation_, sizeof(float) * 16);
  if (result4) {
    result4[0] = g->converged_ ? 1 : 0;
    result4[1] = g->nr_iterations_;
    result4[2] = g->n_linearize_;
    result4[3] = g->n_compute_error_;
  }
  if (final_hessian) std::memcpy(final_hessian, g->final_hessian_, sizeof(double) * 36);
  return std::chrono::duration<double>(t1 - t0).count();
}
double orc_gicp_fitness(void* h, double max_range) { return ((FastGICP*)h)->fitness(max_range); }

// ---- FastVGICP object ----
void* orc_vgicp_create() { return new FastVGICP(); }
void orc_vgicp_destroy(void* h) { delete (FastVGICP*)h; }
void orc_vgicp_set_voxel_params(void* h, double resolution, int search_method, int voxel_mode) {
  FastVGICP* g = (FastVGICP*)h;
  g->voxel_resolution_ = resolution;
  g->search_method_ = search_method;
  g->voxel_mode_ = voxel_mode;
  g->have_voxelmap_ = false;
}
void orc_vgicp_set_target(void* h, const float* xyzw, int n) { ((FastVGICP*)h)->setInputTargetV(xyzw, n); }
double orc_vgicp_linearize(void* h, const double* trans, double* H, double* b) {
  FastVGICP* g = (FastVGICP*)h;
  g->ensure_covariances();
  return g->linearize(trans, H, b);
}
double orc_vgicp_compute_error(void* h, const double* trans) { return ((FastVGICP*)h)->compute_error(trans); }
int orc_vgicp_num_correspondences(void* h) { return (int)((FastVGICP*)h)->voxel_correspondences_.size(); }
// voxels sorted by (x, y, z): coords 3 ints, num, mean 3, cov 6 (upper triangle) per voxel
int orc_vgicp_get_voxels(void* h, int cap, int* coords, int* num, double* mean3, double* cov6) {
  FastVGICP* g = (FastVGICP*)h;
  g->ensure_covariances();
  if (!g->have_voxelmap_) g->create_voxelmap();
  int i = 0;
  for (auto& kv : g->voxels_) {
    if (i >= cap) break;
    coords[3 * i] = kv.first.x; coords[3 * i + 1] = kv.first.y; coords[3 * i + 2] = kv.first.z;
    num[i] = kv.second.num_points;
    for (int d = 0; d < 3; d++) mean3[3 * i + d] = kv.second.mean[d];
    const double* c = kv.second.cov;
    cov6[6 * i] = c[0]; cov6[6 * i + 1] = c[1]; cov6[6 * i + 2] = c[2]; cov6[6 * i + 3] = c[5]; cov6[6 * i + 4] = c[6]; cov6[6 * i + 5] = c[10];
    i++;
  }
  return (int)g->voxels_.size();
}
double orc_vgicp_align(void* h, const float* guess, float* final_T, int* result4) {
  FastVGICP* g = (FastVGICP*)h;
  g->n_linearize_ = g->n_compute_error_ = 0;
  auto t0 = std::chrono::steady_clock::now();
  g->align(guess, nullptr);
  auto t1 = std::chrono::steady_clock::now();
  std::memcpy(final_T, g->final_transformation_, sizeof(float) * 16);
  if (result4) {
    result4[0] = g->converged_ ? 1 : 0;
    result4[1] = g->nr_iterations_;
    result4[2] = g->n_linearize_;
    result4[3] = g->n_compute_error_;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---- A-LOAM feature extraction ----
struct orc_feat_arrays {
  // all arrays caller-allocated with capacity n_in + 8
  float* cloud;  // 4 floats / point
  int* src_index;
  int* intensity_num;
  float *range_vec, *scan_angle, *curvature, *inten_curvature, *curvature2, *distance_source, *other_source;
  int *neighbor_picked, *inten_neighbor_picked, *label, *inten_label, *ground_marked;
  int *corner_sharp, *corner_less_sharp, *surf_flat, *surf_less_flat, *inten_sharp, *inten_less_sharp, *ground_points;
  float *corner_sharp_w, *surf_flat_w, *inten_sharp_w;
  int* scan_start;  // 64
  int* scan_end;    // 64
  int counts[16];   // cloud_size, n_corner_sharp, n_corner_less_sharp, n_surf_flat, n_surf_less_flat, n_inten_sharp,
                    // n_inten_less_sharp, n_ground_points, ground_size, inten_merged
  double groundparam[11];
  double ground_evals[3];
};

double orc_extract_features(const float* xyzi, int n, int n_scans, double min_range, double max_range, int use_intensity, orc_feat_arrays* a) {
  FeatureParams prm;
  prm.n_scans = n_scans;
  prm.minimum_range = min_range;
  prm.maximum_range = max_range;
  prm.use_intensity = use_intensity;
  FeatureOut o;
  auto t0 = std::chrono::steady_clock::now();
  extract_features(xyzi, n, prm, o);
  auto t1 = std::chrono::steady_clock::now();
  if (a) {
    cp(a->cloud, o.cloud);
    cp(a->src_index, o.src_index);
    cp(a->intensity_num, o.intensity_num);
    cp(a->range_vec, o.range_vec);
    cp(a->scan_angle, o.scan_angle);
    cp(a->curvature, o.curvature);
    cp(a->inten_curvature, o.inten_curvature);
    cp(a->curvature2, o.curvature2);
    cp(a->distance_source, o.distance_source);
    cp(a->other_source, o.other_source);
    cp(a->neighbor_picked, o.neighbor_picked);
    cp(a->inten_neighbor_picked, o.inten_neighbor_picked);
    cp(a->label, o.label);
    cp(a->inten_label, o.inten_label);
    cp(a->ground_marked, o.ground_marked);
    cp(a->corner_sharp, o.corner_sharp);
    cp(a->corner_less_sharp, o.corner_less_sharp);
    cp(a->surf_flat, o.surf_flat);
    cp(a->surf_less_flat, o.surf_less_flat);
    cp(a->inten_sharp, o.inten_sharp);
    cp(a->inten_less_sharp, o.inten_less_sharp);
    cp(a->ground_points, o.ground_points);
    cp(a->corner_sharp_w, o.corner_sharp_w);
    cp(a->surf_flat_w, o.surf_flat_w);
    cp(a->inten_sharp_w, o.inten_sharp_w);
    cp(a->scan_start, o.scan_start);
    cp(a->scan_end, o.scan_end);
    int c[16] = {o.cloud_size, (int)o.corner_sharp.size(), (int)o.corner_less_sharp.size(), (int)o.surf_flat.size(), (int)o.surf_less_flat.size(),
                 (int)o.inten_sharp.size(), (int)o.inten_less_sharp.size(), (int)o.ground_points.size(), o.ground_size, o.inten_merged, 0, 0, 0, 0, 0, 0};
    std::memcpy(a->counts, c, sizeof(c));
    std::memcpy(a->groundparam, o.groundparam, sizeof(o.groundparam));
    std::memcpy(a->ground_evals, o.ground_evals, sizeof(o.ground_evals));
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

int orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
