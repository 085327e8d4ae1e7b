// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.hpp header).  PARITY UNPINNED.
//
// CPU restatement of the reference's FastGICP scan-matching path:
//   rgc_slam/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:8-299      (FastGICP)
//   rgc_slam/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:8-172 (LsqRegistration)
//   rgc_slam/include/fast_gicp/so3/so3.hpp:21-77
// plus the parts of pcl::Registration the callers use (align / getFitnessScore), restated from
// their documented behaviour (PCL is un-vendored).  Same OpenMP structure as the reference
// (`parallel for schedule(guided, 8)` over points) so it also serves as the timed CPU baseline
// (cpu_baseline.kind = "port").
//
// Float-arithmetic conventions we had to fix because Eigen's evaluation order cannot be
// inspected here: the float query transform (fast_gicp_impl.hpp:131) is evaluated as
// ((r0*x + r1*y) + r2*z) + t*w with every product and sum rounded to float (no FMA).
#pragma once
#include <omp.h>

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <limits>
#include <map>
#include <vector>

#include "orc_kdtree.hpp"
#include "orc_linalg.hpp"

namespace orc {

enum RegMethod { REG_NONE = 0, REG_MIN_EIG = 1, REG_NORMALIZED_MIN_EIG = 2, REG_PLANE = 3, REG_FROBENIUS = 4 };  // gicp_settings.hpp:6
enum OptType { OPT_GN = 0, OPT_LM = 1 };                                                                           // lsq_registration.hpp:13

struct Mat4d {
  double m[16];  // row-major
};

inline void mat4_mul(const double* A, const double* B, double* C) {
  double T[16];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
      T[i * 4 + j] = s;
    }
  std::memcpy(C, T, sizeof(T));
}

// float Isometry3f * Vector4f (fast_gicp_impl.hpp:131), rows 0..2 only; w passes through.
inline void transform_point_f(const float* T /*row-major 4x4*/, const float* p, float* out) {
  for (int r = 0; r < 3; r++) {
    float a = T[r * 4 + 0] * p[0];
    float b = T[r * 4 + 1] * p[1];
    float c = T[r * 4 + 2] * p[2];
    float d = T[r * 4 + 3] * p[3];
    float s = a + b;
    s = s + c;
    s = s + d;
    out[r] = s;
  }
  out[3] = p[3];
}

class FastGICP {
 public:
  // ---- parameters (defaults: fast_gicp_impl.hpp:8-23, lsq_registration_impl.hpp:9-22) ----
  int num_threads_ = omp_get_max_threads();
  int k_correspondences_ = 20;
  float corr_dist_threshold_ = std::numeric_limits<float>::max();
  int regularization_method_ = REG_PLANE;
  int max_iterations_ = 64;
  double rotation_epsilon_ = 2e-3;
  double transformation_epsilon_ = 5e-4;
  int lsq_optimizer_type_ = OPT_LM;
  int lm_max_iterations_ = 10;
  double lm_init_lambda_factor_ = 1e-9;
  double lm_lambda_ = -1.0;
  bool lm_debug_print_ = false;

  // ---- state ----
  std::vector<float> input_, target_;  // xyzw float4 per point
  KdTree source_kdtree_, target_kdtree_;
  std::vector<Mat4d> source_covs_, target_covs_, mahalanobis_;
  std::vector<int> correspondences_;
  std::vector<float> sq_distances_;
  double final_hessian_[36];
  float final_transformation_[16];
  bool converged_ = false;
  int nr_iterations_ = 0;
  int n_linearize_ = 0, n_compute_error_ = 0;

  FastGICP() {
    for (int i = 0; i < 36; i++) final_hessian_[i] = (i % 7 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 16; i++) final_transformation_[i] = (i % 5 == 0) ? 1.f : 0.f;
  }

  int n_source() const { return (int)(input_.size() / 4); }
  int n_target() const { return (int)(target_.size() / 4); }

  // fast_gicp_impl.hpp:72-81 (identity caching is the caller's business in this flat API)
  void setInputSource(const float* xyzw, int n) {
    input_.assign(xyzw, xyzw + 4 * (size_t)n);
    source_kdtree_.build(input_.data(), n);
    source_covs_.clear();
  }
  // fast_gicp_impl.hpp:83-91
  void setInputTarget(const float* xyzw, int n) {
    target_.assign(xyzw, xyzw + 4 * (size_t)n);
    target_kdtree_.build(target_.data(), n);
    target_covs_.clear();
  }

  // fast_gicp_impl.hpp:241-299
  void calculate_covariances(const std::vector<float>& cloud, const KdTree& kdtree, std::vector<Mat4d>& covariances) const {
    const int n = (int)(cloud.size() / 4);
    const int k = k_correspondences_;
    covariances.resize(n);
#pragma omp parallel for num_threads(num_threads_) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      std::vector<int> k_indices(k);
      std::vector<float> k_sq(k);
      int found = kdtree.knn(&cloud[4 * (size_t)i], k, k_indices.data(), k_sq.data());
      covariance_from_neighbors(cloud.data(), k_indices.data(), found, k, regularization_method_, covariances[i].m);
    }
  }

  // :256-293 — `neighbors` is 4 x k with only `found` columns filled; Eigen leaves the remaining
  // columns uninitialised in the reference (k > N is outside its contract) — we zero them.
  static void covariance_from_neighbors(const float* cloud, const int* idx, int found, int k, int method, double* out16) {
    std::vector<double> nb(4 * (size_t)k, 0.0);
    for (int j = 0; j < found; j++)
      for (int d = 0; d < 4; d++) nb[d * (size_t)k + j] = (double)cloud[4 * (size_t)idx[j] + d];
    for (int d = 0; d < 4; d++) {
      double mean = 0.0;
      for (int j = 0; j < k; j++) mean += nb[d * (size_t)k + j];
      mean /= k;
      for (int j = 0; j < k; j++) nb[d * (size_t)k + j] -= mean;
    }
    double cov[16];
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++) {
        double s = 0.0;
        for (int j = 0; j < k; j++) s += nb[a * (size_t)k + j] * nb[b * (size_t)k + j];
        cov[a * 4 + b] = s / k;
      }
    if (method == REG_NONE) {
      std::memcpy(out16, cov, sizeof(cov));
      return;
    }
    double C3[9];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) C3[a * 3 + b] = cov[a * 4 + b];
    double R3[9];
    if (method == REG_FROBENIUS) {
      const double lambda = 1e-3;
      double C[9], Ci[9];
      for (int i = 0; i < 9; i++) C[i] = C3[i] + ((i % 4 == 0) ? lambda : 0.0);
      inverse3(C, Ci);
      double nrm = 0.0;
      for (int i = 0; i < 9; i++) nrm += Ci[i] * Ci[i];
      nrm = std::sqrt(nrm);
      for (int i = 0; i < 9; i++) Ci[i] /= nrm;
      inverse3(Ci, R3);
    } else {
      double U[9], sv[3], V[9], values[3];
      jacobi_svd3(C3, U, sv, V);
      switch (method) {
        case REG_PLANE:
          values[0] = 1.0; values[1] = 1.0; values[2] = 1e-3;
          break;
        case REG_MIN_EIG:
          for (int i = 0; i < 3; i++) values[i] = std::max(sv[i], 1e-3);
          break;
        case REG_NORMALIZED_MIN_EIG: {
          double mx = std::max(sv[0], std::max(sv[1], sv[2]));
          for (int i = 0; i < 3; i++) values[i] = std::max(sv[i] / mx, 1e-3);
          break;
        }
        default:
          std::fprintf(stderr, "here must not be reached\n");
          std::abort();
      }
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          double s = 0.0;
          for (int c = 0; c < 3; c++) s += U[a * 3 + c] * values[c] * V[b * 3 + c];
          R3[a * 3 + b] = s;
        }
    }
    for (int i = 0; i < 16; i++) out16[i] = 0.0;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) out16[a * 4 + b] = R3[a * 3 + b];
  }

  void ensure_covariances() {  // fast_gicp_impl.hpp:103-109
    if ((int)source_covs_.size() != n_source()) calculate_covariances(input_, source_kdtree_, source_covs_);
    if ((int)target_covs_.size() != n_target()) calculate_covariances(target_, target_kdtree_, target_covs_);
  }

  // fast_gicp_impl.hpp:115-152
  void update_correspondences(const double* trans /*row-major 4x4*/) {
    const int n = n_source();
    float trans_f[16];
    for (int i = 0; i < 16; i++) trans_f[i] = (float)trans[i];
    correspondences_.resize(n);
    sq_distances_.resize(n);
    mahalanobis_.resize(n);
    const float thr2 = corr_dist_threshold_ * corr_dist_threshold_;
#pragma omp parallel for num_threads(num_threads_) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      float pt[4];
      transform_point_f(trans_f, &input_[4 * (size_t)i], pt);
      int k_index = -1;
      float k_sq = 0.f;
      target_kdtree_.knn(pt, 1, &k_index, &k_sq);
      sq_distances_[i] = k_sq;
      correspondences_[i] = k_sq < thr2 ? k_index : -1;
      if (correspondences_[i] < 0) continue;
      const double* cov_A = source_covs_[i].m;
      const double* cov_B = target_covs_[correspondences_[i]].m;
      double TC[16], Tt[16], RCR[16];
      mat4_mul(trans, cov_A, TC);
      for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) Tt[a * 4 + b] = trans[b * 4 + a];
      mat4_mul(TC, Tt, RCR);
      for (int j = 0; j < 16; j++) RCR[j] = cov_B[j] + RCR[j];
      RCR[15] = 1.0;
      inverse4(RCR, mahalanobis_[i].m);
      mahalanobis_[i].m[15] = 0.0;
    }
  }

  // fast_gicp_impl.hpp:155-211 ; H row-major 6x6, b 6.  H/b may be null.
  double linearize(const double* trans, double* H, double* b) {
    n_linearize_++;
    update_correspondences(trans);
    const int n = n_source();
    double sum_errors = 0.0;
    std::vector<double> Hs((size_t)num_threads_ * 36, 0.0), bs((size_t)num_threads_ * 6, 0.0);
#pragma omp parallel for num_threads(num_threads_) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      int ti = correspondences_[i];
      if (ti < 0) continue;
      double mean_A[4], mean_B[4], tA[4], err[4];
      for (int d = 0; d < 4; d++) {
        mean_A[d] = (double)input_[4 * (size_t)i + d];
        mean_B[d] = (double)target_[4 * (size_t)ti + d];
      }
      for (int r = 0; r < 3; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += trans[r * 4 + c] * mean_A[c];
        tA[r] = s;
      }
      tA[3] = mean_A[3];
      for (int d = 0; d < 4; d++) err[d] = mean_B[d] - tA[d];
      const double* M = mahalanobis_[i].m;
      double Me[4];
      for (int r = 0; r < 4; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += M[r * 4 + c] * err[c];
        Me[r] = s;
      }
      sum_errors += err[0] * Me[0] + err[1] * Me[1] + err[2] * Me[2] + err[3] * Me[3];
      if (H == nullptr || b == nullptr) continue;
      double J[24] = {0};  // 4 x 6
      // skewd(tA.head<3>()) (so3.hpp:21-31) and -I
      J[0 * 6 + 1] = -tA[2]; J[0 * 6 + 2] = tA[1];
      J[1 * 6 + 0] = tA[2];  J[1 * 6 + 2] = -tA[0];
      J[2 * 6 + 0] = -tA[1]; J[2 * 6 + 1] = tA[0];
      J[0 * 6 + 3] = -1.0; J[1 * 6 + 4] = -1.0; J[2 * 6 + 5] = -1.0;
      double MJ[24];
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 6; c++) {
          double s = 0.0;
          for (int k = 0; k < 4; k++) s += M[r * 4 + k] * J[k * 6 + c];
          MJ[r * 6 + c] = s;
        }
      double* Ht = &Hs[(size_t)omp_get_thread_num() * 36];
      double* bt = &bs[(size_t)omp_get_thread_num() * 6];
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) {
          double s = 0.0;
          for (int k = 0; k < 4; k++) s += J[k * 6 + r] * MJ[k * 6 + c];
          Ht[r * 6 + c] += s;
        }
        double s = 0.0;
        for (int k = 0; k < 4; k++) s += J[k * 6 + r] * Me[k];
        bt[r] += s;
      }
    }
    if (H && b) {
      for (int i = 0; i < 36; i++) H[i] = 0.0;
      for (int i = 0; i < 6; i++) b[i] = 0.0;
      for (int t = 0; t < num_threads_; t++) {
        for (int i = 0; i < 36; i++) H[i] += Hs[(size_t)t * 36 + i];
        for (int i = 0; i < 6; i++) b[i] += bs[(size_t)t * 6 + i];
      }
    }
    return sum_errors;
  }

  // fast_gicp_impl.hpp:214-237 — reuses stale correspondences_ and mahalanobis_ on purpose.
  double compute_error(const double* trans) {
    n_compute_error_++;
    const int n = n_source();
    double sum_errors = 0.0;
#pragma omp parallel for num_threads(num_threads_) reduction(+ : sum_errors) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      int ti = correspondences_[i];
      if (ti < 0) continue;
      double err[4];
      for (int r = 0; r < 3; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += trans[r * 4 + c] * (double)input_[4 * (size_t)i + c];
        err[r] = (double)target_[4 * (size_t)ti + r] - s;
      }
      err[3] = (double)target_[4 * (size_t)ti + 3] - (double)input_[4 * (size_t)i + 3];
      const double* M = mahalanobis_[i].m;
      double acc = 0.0;
      for (int r = 0; r < 4; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += M[r * 4 + c] * err[c];
        acc += err[r] * s;
      }
      sum_errors += acc;
    }
    return sum_errors;
  }

  // lsq_registration_impl.hpp:82-91
  bool is_converged(const double* delta) const {
    double rmax = 0.0, tmax = 0.0;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) {
        double v = std::fabs(delta[r * 4 + c] - (r == c ? 1.0 : 0.0));
        rmax = std::max(rmax, 1.0 / rotation_epsilon_ * v);
      }
      tmax = std::max(tmax, 1.0 / transformation_epsilon_ * std::fabs(delta[r * 4 + 3]));
    }
    return std::max(rmax, tmax) < 1;
  }

  static void make_delta(const double* d, double* delta) {
    double q[4], R[9];
    so3_exp_quat(d, q);
    quat_to_rot(q, R);
    for (int i = 0; i < 16; i++) delta[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) delta[r * 4 + c] = R[r * 3 + c];
      delta[r * 4 + 3] = d[3 + r];
    }
  }

  // lsq_registration_impl.hpp:106-122
  bool step_gn(double* x0, double* delta) {
    double H[36], b[6], nb[6], d[6];
    linearize(x0, H, b);
    for (int i = 0; i < 6; i++) nb[i] = -b[i];
    ldlt6_solve(H, nb, d);
    make_delta(d, delta);
    mat4_mul(delta, x0, x0);
    std::memcpy(final_hessian_, H, sizeof(H));
    return true;
  }

  // lsq_registration_impl.hpp:125-172
  bool step_lm(double* x0, double* delta) {
    double H[36], b[6];
    double y0 = linearize(x0, H, b);
    if (lm_lambda_ < 0.0) {
      double mx = 0.0;
      for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i * 6 + i]));
      lm_lambda_ = lm_init_lambda_factor_ * mx;
    }
    double nu = 2.0;
    for (int i = 0; i < lm_max_iterations_; i++) {
      double A[36], nb[6], d[6];
      for (int j = 0; j < 36; j++) A[j] = H[j] + ((j % 7 == 0) ? lm_lambda_ : 0.0);
      for (int j = 0; j < 6; j++) nb[j] = -b[j];
      ldlt6_solve(A, nb, d);
      make_delta(d, delta);
      double xi[16];
      mat4_mul(delta, x0, xi);
      double yi = compute_error(xi);
      double denom = 0.0;
      for (int j = 0; j < 6; j++) denom += d[j] * (lm_lambda_ * d[j] - b[j]);
      double rho = (y0 - yi) / denom;
      if (lm_debug_print_) std::printf("%5d %15g %15g %15g %15g\n", i, y0, yi, rho, lm_lambda_);
      if (rho < 0) {
        if (is_converged(delta)) return true;
        lm_lambda_ = nu * lm_lambda_;
        nu = 2 * nu;
        continue;
      }
      std::memcpy(x0, xi, sizeof(xi));
      lm_lambda_ = lm_lambda_ * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(final_hessian_, H, sizeof(H));
      return true;
    }
    return false;
  }

  // pcl::Registration::align(output, guess) -> FastGICP::computeTransformation
  // (fast_gicp_impl.hpp:103-112) -> LsqRegistration::computeTransformation
  // (lsq_registration_impl.hpp:53-79).  guess/result row-major float 4x4.
  void align(const float* guess, float* out_points /*nullable, 4 floats per source point*/) {
    ensure_covariances();
    double x0[16];
    for (int i = 0; i < 16; i++) x0[i] = (double)guess[i];
    lm_lambda_ = -1.0;
    converged_ = false;
    nr_iterations_ = 0;
    for (int i = 0; i < max_iterations_ && !converged_; i++) {
      nr_iterations_ = i;
      double delta[16];
      bool ok = (lsq_optimizer_type_ == OPT_GN) ? step_gn(x0, delta) : step_lm(x0, delta);
      if (!ok) {
        std::fprintf(stderr, "lm not converged!!\n");
        break;
      }
      converged_ = is_converged(delta);
    }
    for (int i = 0; i < 16; i++) final_transformation_[i] = (float)x0[i];
    if (out_points) transform_cloud(final_transformation_, out_points);
  }

  // pcl::transformPointCloud(*input_, output, final_transformation_) (lsq_registration_impl.hpp:78)
  void transform_cloud(const float* T, float* out_points) const {
    const int n = n_source();
    for (int i = 0; i < n; i++) {
      const float* p = &input_[4 * (size_t)i];
      float one[4] = {p[0], p[1], p[2], 1.f};
      transform_point_f(T, one, &out_points[4 * (size_t)i]);
      out_points[4 * (size_t)i + 3] = 1.f;
    }
  }

  // pcl::Registration::getFitnessScore(max_range): mean of 1-NN squared distances of the
  // transformed source, keeping d2 <= max_range; DBL_MAX if none.  (PCL un-vendored; callers
  // RGC_odometer.cpp:1010, RGC_mapping.cpp:2070.)  Single-threaded in PCL; parallel here with
  // an ordered (serial) final sum so the value is deterministic.
  double fitness(double max_range) const {
    const int n = n_source();
    std::vector<float> d2(n);
#pragma omp parallel for num_threads(num_threads_) schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      const float* p = &input_[4 * (size_t)i];
      float one[4] = {p[0], p[1], p[2], 1.f}, q[4];
      transform_point_f(final_transformation_, one, q);
      int id;
      float d;
      target_kdtree_.knn(q, 1, &id, &d);
      d2[i] = d;
    }
    double s = 0.0;
    int nr = 0;
    for (int i = 0; i < n; i++)
      if ((double)d2[i] <= max_range) {
        s += (double)d2[i];
        nr++;
      }
    return nr > 0 ? s / nr : std::numeric_limits<double>::max();
  }
};


// ------------------------------------------------------------------------------------------------
// FastVGICP (what RGC_odometer.cpp:998 actually instantiates) — SURVEY.md §8f N1.
//   rgc_slam/include/fast_gicp/gicp/impl/fast_vgicp_impl.hpp:73-204
//   rgc_slam/include/fast_gicp/gicp/fast_vgicp_voxel.hpp:10-183
enum NeighborSearch { DIRECT27 = 0, DIRECT7 = 1, DIRECT1 = 2 };                  // gicp_settings.hpp:8
enum VoxelMode { ADDITIVE = 0, ADDITIVE_WEIGHTED = 1, MULTIPLICATIVE = 2 };      // gicp_settings.hpp:10

struct GaussianVoxel {
  int num_points = 0;
  double mean[4] = {0, 0, 0, 0};
  double cov[16] = {0};
};

class FastVGICP : public FastGICP {
 public:
  double voxel_resolution_ = 1.0;   // fast_vgicp_impl.hpp:20
  int search_method_ = DIRECT1;     // :21
  int voxel_mode_ = ADDITIVE;       // :22
  struct Key {
    int x, y, z;
    bool operator<(const Key& o) const { return x != o.x ? x < o.x : (y != o.y ? y < o.y : z < o.z); }
  };
  std::map<Key, GaussianVoxel> voxels_;
  bool have_voxelmap_ = false;
  std::vector<std::pair<int, const GaussianVoxel*>> voxel_correspondences_;
  std::vector<Mat4d> voxel_mahalanobis_;

  // fast_vgicp_voxel.hpp:158-160
  Key voxel_coord(const double* x) const {
    return Key{(int)std::floor(x[0] / voxel_resolution_ - 0.5), (int)std::floor(x[1] / voxel_resolution_ - 0.5),
               (int)std::floor(x[2] / voxel_resolution_ - 0.5)};
  }
  static std::vector<Key> neighbor_offsets(int method) {  // fast_vgicp_voxel.hpp:10-44
    if (method == DIRECT1) return {{0, 0, 0}};
    if (method == DIRECT7) return {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    std::vector<Key> o;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) o.push_back(Key{i - 1, j - 1, k - 1});
    return o;
  }
  // fast_vgicp_voxel.hpp:129-156 (sequential, points appended in index order)
  void create_voxelmap() {
    voxels_.clear();
    const int n = n_target();
    for (int i = 0; i < n; i++) {
      double m[4];
      for (int d = 0; d < 4; d++) m[d] = (double)target_[4 * (size_t)i + d];
      GaussianVoxel& v = voxels_[voxel_coord(m)];
      const double* c = target_covs_[i].m;
      v.num_points++;
      if (voxel_mode_ == MULTIPLICATIVE) {  // :86-93
        double ci[16], inv[16];
        std::memcpy(ci, c, sizeof(ci));
        ci[15] = 1;
        inverse4(ci, inv);
        for (int j = 0; j < 16; j++) v.cov[j] += inv[j];
        for (int r = 0; r < 4; r++)
          for (int cc = 0; cc < 4; cc++) v.mean[r] += inv[r * 4 + cc] * m[cc];
      } else {  // :110-114
        for (int d = 0; d < 4; d++) v.mean[d] += m[d];
        for (int j = 0; j < 16; j++) v.cov[j] += c[j];
      }
    }
    for (auto& kv : voxels_) {
      GaussianVoxel& v = kv.second;
      if (voxel_mode_ == MULTIPLICATIVE) {  // :95-100
        v.cov[15] = 1;
        v.mean[3] = 1;
        double inv[16], mm[4];
        inverse4(v.cov, inv);
        std::memcpy(v.cov, inv, sizeof(inv));
        for (int r = 0; r < 4; r++) {
          mm[r] = 0;
          for (int cc = 0; cc < 4; cc++) mm[r] += v.cov[r * 4 + cc] * v.mean[cc];
        }
        std::memcpy(v.mean, mm, sizeof(mm));
      } else {  // :116-119
        for (int d = 0; d < 4; d++) v.mean[d] /= v.num_points;
        for (int j = 0; j < 16; j++) v.cov[j] /= v.num_points;
      }
    }
    have_voxelmap_ = true;
  }
  void setInputTargetV(const float* xyzw, int n) {
    setInputTarget(xyzw, n);
    have_voxelmap_ = false;
  }

  // fast_vgicp_impl.hpp:73-116 (correspondences kept in (point, offset) order: the reference's order
  // depends on the OpenMP schedule)
  void update_correspondences_v(const double* trans) {
    voxel_correspondences_.clear();
    const auto offsets = neighbor_offsets(search_method_);
    const int n = n_source();
    for (int i = 0; i < n; i++) {
      double tA[4];
      for (int r = 0; r < 3; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += trans[r * 4 + c] * (double)input_[4 * (size_t)i + c];
        tA[r] = s;
      }
      tA[3] = 1.0;
      const Key coord = voxel_coord(tA);
      for (const Key& o : offsets) {
        auto it = voxels_.find(Key{coord.x + o.x, coord.y + o.y, coord.z + o.z});
        if (it != voxels_.end()) voxel_correspondences_.push_back({i, &it->second});
      }
    }
    voxel_mahalanobis_.resize(voxel_correspondences_.size());
#pragma omp parallel for num_threads(num_threads_) schedule(guided, 8)
    for (int i = 0; i < (int)voxel_correspondences_.size(); i++) {
      const double* cov_A = source_covs_[voxel_correspondences_[i].first].m;
      const double* cov_B = voxel_correspondences_[i].second->cov;
      double TC[16], Tt[16], RCR[16];
      mat4_mul(trans, cov_A, TC);
      for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) Tt[a * 4 + b] = trans[b * 4 + a];
      mat4_mul(TC, Tt, RCR);
      for (int j = 0; j < 16; j++) RCR[j] = cov_B[j] + RCR[j];
      RCR[15] = 1.0;
      inverse4(RCR, voxel_mahalanobis_[i].m);
      voxel_mahalanobis_[i].m[15] = 0.0;
    }
  }

  double accumulate_v(const double* trans, double* H, double* b) {
    double sum_errors = 0.0;
    if (H && b) {
      for (int i = 0; i < 36; i++) H[i] = 0.0;
      for (int i = 0; i < 6; i++) b[i] = 0.0;
    }
    for (size_t ci = 0; ci < voxel_correspondences_.size(); ci++) {
      const int i = voxel_correspondences_[ci].first;
      const GaussianVoxel* vx = voxel_correspondences_[ci].second;
      double tA[4], err[4];
      for (int r = 0; r < 3; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += trans[r * 4 + c] * (double)input_[4 * (size_t)i + c];
        tA[r] = s;
      }
      tA[3] = (double)input_[4 * (size_t)i + 3];
      for (int d = 0; d < 4; d++) err[d] = vx->mean[d] - tA[d];
      const double* M = voxel_mahalanobis_[ci].m;
      const double w = std::sqrt((double)vx->num_points);
      double Me[4];
      for (int r = 0; r < 4; r++) {
        double s = 0.0;
        for (int c = 0; c < 4; c++) s += M[r * 4 + c] * err[c];
        Me[r] = s;
      }
      sum_errors += w * (err[0] * Me[0] + err[1] * Me[1] + err[2] * Me[2] + err[3] * Me[3]);
      if (!H || !b) continue;
      double J[24] = {0};
      J[0 * 6 + 1] = -tA[2]; J[0 * 6 + 2] = tA[1];
      J[1 * 6 + 0] = tA[2];  J[1 * 6 + 2] = -tA[0];
      J[2 * 6 + 0] = -tA[1]; J[2 * 6 + 1] = tA[0];
      J[0 * 6 + 3] = -1.0; J[1 * 6 + 4] = -1.0; J[2 * 6 + 5] = -1.0;
      double MJ[24];
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 6; c++) {
          double s = 0.0;
          for (int k = 0; k < 4; k++) s += M[r * 4 + k] * J[k * 6 + c];
          MJ[r * 6 + c] = s;
        }
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) {
          double s = 0.0;
          for (int k = 0; k < 4; k++) s += J[k * 6 + r] * MJ[k * 6 + c];
          H[r * 6 + c] += w * s;
        }
        double s = 0.0;
        for (int k = 0; k < 4; k++) s += J[k * 6 + r] * Me[k];
        b[r] += w * s;
      }
    }
    return sum_errors;
  }
  // fast_vgicp_impl.hpp:119-180
  double linearize(const double* trans, double* H, double* b) {
    n_linearize_++;
    if (!have_voxelmap_) create_voxelmap();
    update_correspondences_v(trans);
    return accumulate_v(trans, H, b);
  }
  // :183-204
  double compute_error(const double* trans) {
    n_compute_error_++;
    return accumulate_v(trans, nullptr, nullptr);
  }
  // lsq_registration_impl.hpp:106-172 with the voxelised linearize / compute_error
  bool step_lm_v(double* x0, double* delta) {
    double H[36], b[6];
    double y0 = linearize(x0, H, b);
    if (lm_lambda_ < 0.0) {
      double mx = 0.0;
      for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i * 6 + i]));
      lm_lambda_ = lm_init_lambda_factor_ * mx;
    }
    double nu = 2.0;
    for (int i = 0; i < lm_max_iterations_; i++) {
      double A[36], nb[6], d[6], xi[16];
      for (int j = 0; j < 36; j++) A[j] = H[j] + ((j % 7 == 0) ? lm_lambda_ : 0.0);
      for (int j = 0; j < 6; j++) nb[j] = -b[j];
      ldlt6_solve(A, nb, d);
      make_delta(d, delta);
      mat4_mul(delta, x0, xi);
      double yi = compute_error(xi);
      double denom = 0.0;
      for (int j = 0; j < 6; j++) denom += d[j] * (lm_lambda_ * d[j] - b[j]);
      double rho = (y0 - yi) / denom;
      if (rho < 0) {
        if (is_converged(delta)) return true;
        lm_lambda_ = nu * lm_lambda_;
        nu = 2 * nu;
        continue;
      }
      std::memcpy(x0, xi, sizeof(xi));
      lm_lambda_ = lm_lambda_ * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(final_hessian_, H, sizeof(H));
      return true;
    }
    return false;
  }
  void align(const float* guess, float* out_points) {
    have_voxelmap_ = false;  // computeTransformation resets the voxel map (fast_vgicp_impl.hpp:66-70)
    ensure_covariances();
    double x0[16];
    for (int i = 0; i < 16; i++) x0[i] = (double)guess[i];
    lm_lambda_ = -1.0;
    converged_ = false;
    nr_iterations_ = 0;
    for (int i = 0; i < max_iterations_ && !converged_; i++) {
      nr_iterations_ = i;
      double delta[16];
      if (!step_lm_v(x0, delta)) {
        std::fprintf(stderr, "lm not converged!!\n");
        break;
      }
      converged_ = is_converged(delta);
    }
    for (int i = 0; i < 16; i++) final_transformation_[i] = (float)x0[i];
    if (out_points) transform_cloud(final_transformation_, out_points);
  }
};

}  // namespace orc
