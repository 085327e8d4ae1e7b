// rgc/fast_gicp.hpp — header-only C++ facade over the C-ABI (include/rgc_gicp.h).
//
// Keeps the pcl::Registration-style API the reference's odometry and mapping nodes call
// (/root/reference/rgc_slam/src/RGC_odometer.cpp:998-1011):
//     rgc::FastGICP<PointT, PointT> gicp;
//     gicp.setMaximumIterations(25); gicp.setMaxCorrespondenceDistance(2);
//     gicp.setTransformationEpsilon(1e-6); gicp.setNumThreads(14);
//     gicp.setInputTarget(target); gicp.setInputSource(source);
//     gicp.align(aligned, guess);
//     gicp.getFitnessScore(); gicp.getFinalTransformation();
// mirroring fast_gicp::FastGICP (rgc_slam/include/fast_gicp/gicp/fast_gicp.hpp:20-100) and
// fast_gicp::LsqRegistration (lsq_registration.hpp:16-86).
//
// Two flavours, chosen at compile time:
//  * with PCL on the include path (#define RGC_WITH_PCL or auto-detected) the class derives from
//    pcl::Registration<PointSource, PointTarget, float> and overrides the same three virtuals the
//    reference overrides (setInputSource, setInputTarget, computeTransformation) — a true drop-in;
//  * without PCL (this repository's build image has none) it is a stand-alone class with the same
//    member names over the minimal rgc::PointCloud below, Matrix4f being a column-major
//    std::array<float,16> (Eigen::Matrix4f::data() layout).
//
// The object is as cheap to construct per frame as the reference's stack-local one: the CUDA
// stream, pooled device memory and pinned buffers live in a process-wide rgc_ctx per device.
#pragma once
#include <array>
#include <cfloat>
#include <cstdint>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../rgc_gicp.h"
#include "../rgc_preprocess.h"

#if !defined(RGC_WITH_PCL) && defined(__has_include)
#if __has_include(<pcl/registration/registration.h>)
#define RGC_WITH_PCL 1
#endif
#endif

namespace rgc {

enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS };  // gicp_settings.hpp:6
enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };                        // lsq_registration.hpp:13

// One context per (host thread, device): the caller constructs a registration object every frame
// (RGC_odometer.cpp:998), so nothing expensive may live in the object itself, and a context (streams,
// memory pool, pinned result buffers) must not be shared by threads running concurrently.  Threads
// that each run their own registrations (batched loop-closure verification) overlap on the GPU.
// The context is destroyed when its thread ends: registration objects are meant to be locals (as at the
// reference's call site) and must not outlive the thread that created them.
inline rgc_ctx* shared_context(int device = 0) {
  struct Holder {
    std::vector<rgc_ctx*> ctxs;
    ~Holder() {
      for (rgc_ctx* c : ctxs)
        if (c) rgc_ctx_destroy(c);
    }
  };
  static thread_local Holder h;
  if ((int)h.ctxs.size() <= device) h.ctxs.resize(device + 1, nullptr);
  if (!h.ctxs[device]) {
    if (rgc_ctx_create(device, &h.ctxs[device]) != RGC_OK)
      throw std::runtime_error("rgc: no usable CUDA device " + std::to_string(device) + " (there is no CPU fallback)");
  }
  return h.ctxs[device];
}

namespace detail {
inline void check(rgc_ctx* ctx, int rc) {
  if (rc != RGC_OK) throw std::runtime_error(std::string("rgc: ") + rgc_last_error(ctx));
}

// everything that talks to the C-ABI, shared by both flavours
class Core {
 public:
  explicit Core(int device = 0) : ctx_(shared_context(device)) {
    check(ctx_, rgc_reg_create(ctx_, &reg_));
    rgc_params_default(&prm_);
  }
  ~Core() { rgc_reg_destroy(reg_); }
  Core(const Core&) = delete;
  Core& operator=(const Core&) = delete;

  void push() { check(ctx_, rgc_reg_set_params(reg_, &prm_)); }
  void setSource(const void* pts, size_t n, size_t stride, uint64_t key) { check(ctx_, rgc_reg_set_source(reg_, pts, n, stride, key)); n_src_ = n; }
  void setTarget(const void* pts, size_t n, size_t stride, uint64_t key) { check(ctx_, rgc_reg_set_target(reg_, pts, n, stride, key)); n_tgt_ = n; }
  size_t setFiltered(bool source, const void* pts, size_t n, size_t stride, size_t inten_off, float leaf, const double* q_wxyz, const double* t3,
                     float scan_period, uint64_t key) {
    size_t m = 0;
    check(ctx_, (source ? rgc_reg_set_source_filtered : rgc_reg_set_target_filtered)(reg_, pts, n, stride, inten_off, leaf, q_wxyz, t3, scan_period, key, &m));
    (source ? n_src_ : n_tgt_) = m;
    return m;
  }
  void align(const float* guess16, float* final16, float* out_points) {
    check(ctx_, rgc_reg_align(reg_, guess16, final16, &res_, out_points));
  }
  double fitness(double max_range) {
    double s = 0;
    check(ctx_, rgc_reg_fitness(reg_, max_range, &s));
    return s;
  }
  double linearize(const double* T16, double* H36, double* b6) {
    double e = 0;
    check(ctx_, rgc_reg_linearize(reg_, T16, &e, H36, b6));
    return e;
  }
  rgc_ctx* ctx_;
  rgc_reg* reg_ = nullptr;
  rgc_params prm_;
  rgc_result res_{};
  size_t n_src_ = 0, n_tgt_ = 0;
};

// byte offset of PointT::intensity, or RGC_NO_INTENSITY for point types without one
template <class P, class = void>
struct IntensityOffset {
  static size_t get() { return RGC_NO_INTENSITY; }
};
template <class P>
struct IntensityOffset<P, decltype((void)std::declval<P&>().intensity)> {
  static size_t get() {
    P p{};
    return (size_t)(reinterpret_cast<const char*>(&p.intensity) - reinterpret_cast<const char*>(&p));
  }
};
// swapSourceAndTarget on the host-side handles: same pointer type -> swap, different -> both reset
template <class A>
inline void swap_handles(A& a, A& b) {
  std::swap(a, b);
}
template <class A, class B>
inline void swap_handles(A& a, B& b) {
  a.reset();
  b.reset();
}
}  // namespace detail

#if !defined(RGC_WITH_PCL)
// ---- minimal stand-ins for the PCL types the API mentions ---------------------------------------
struct alignas(16) PointXYZ {
  float x, y, z, w = 1.f;
};
struct alignas(16) PointXYZI {
  float x, y, z, w = 1.f;
  float intensity = 0.f, pad[3] = {0, 0, 0};
};
template <class PointT>
struct PointCloud {
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  size_t size() const { return points.size(); }
  void resize(size_t n) { points.resize(n); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
};
using Matrix4f = std::array<float, 16>;  // column-major
inline Matrix4f identity4() { return Matrix4f{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

template <class PointSource, class PointTarget>
class FastGICP {
 public:
  using PointCloudSource = PointCloud<PointSource>;
  using PointCloudTarget = PointCloud<PointTarget>;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Matrix4 = Matrix4f;

  explicit FastGICP(int device = 0) : core_(device) {}

  // ---- pcl::Registration setters used by the reference's callers ----
  void setMaximumIterations(int n) { core_.prm_.max_iterations = n; core_.push(); }
  void setMaxCorrespondenceDistance(double d) { core_.prm_.max_correspondence_distance = (float)d; core_.push(); }
  void setTransformationEpsilon(double e) { core_.prm_.transformation_epsilon = e; core_.push(); }
  void setEuclideanFitnessEpsilon(double) {}  // no-op for LsqRegistration-based classes
  void setRANSACIterations(int) {}            // no-op
  // ---- fast_gicp extras (fast_gicp.hpp:51-72, lsq_registration.hpp:51-61) ----
  void setNumThreads(int) {}  // OpenMP knob: meaningless on the GPU, accepted and ignored
  void setCorrespondenceRandomness(int k) { core_.prm_.k_correspondences = k; core_.push(); }
  void setRegularizationMethod(RegularizationMethod m) { core_.prm_.regularization = (int)m; core_.push(); }
  void setRotationEpsilon(double e) { core_.prm_.rotation_epsilon = e; core_.push(); }
  void setInitialLambdaFactor(double f) { core_.prm_.lm_init_lambda_factor = f; core_.push(); }
  void setDebugPrint(bool on) { core_.prm_.lm_debug_print = on ? 1 : 0; core_.push(); }
  void setLSQType(LSQ_OPTIMIZER_TYPE t) { core_.prm_.optimizer = (int)t; core_.push(); }

  // shared_ptr identity decides whether anything is recomputed (fast_gicp_impl.hpp:72-91)
  void setInputSource(const PointCloudSourceConstPtr& cloud) {
    source_ = cloud;
    filtered_source_ = false;
    core_.setSource(cloud->points.data(), cloud->size(), sizeof(PointSource), (uint64_t)(uintptr_t)cloud.get());
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    target_ = cloud;
    filtered_target_ = false;
    core_.setTarget(cloud->points.data(), cloud->size(), sizeof(PointTarget), (uint64_t)(uintptr_t)cloud.get());
  }
  // The frame's front end fused into setInput* (include/rgc_preprocess.h): [adjustDistortion with
  // q_last_curr (w, x, y, z) / t_last_curr, RGC_odometer.cpp:1441-1481] -> pcl::VoxelGrid(leaf)
  // (:975-991) -> setInputSource / setInputTarget, without the cloud leaving the device.  Returns
  // the size of the filtered cloud.  align()'s output cloud then has that many points.
  size_t setInputSourceFiltered(const PointCloudSourceConstPtr& cloud, float leaf, const double* q_last_curr_wxyz = nullptr,
                                const double* t_last_curr = nullptr, float scan_period = 0.1f) {
    source_ = cloud;
    filtered_source_ = true;
    return core_.setFiltered(true, cloud->points.data(), cloud->size(), sizeof(PointSource), detail::IntensityOffset<PointSource>::get(), leaf,
                             q_last_curr_wxyz, t_last_curr, scan_period, (uint64_t)(uintptr_t)cloud.get());
  }
  size_t setInputTargetFiltered(const PointCloudTargetConstPtr& cloud, float leaf, const double* q_last_curr_wxyz = nullptr,
                                const double* t_last_curr = nullptr, float scan_period = 0.1f) {
    target_ = cloud;
    filtered_target_ = true;
    return core_.setFiltered(false, cloud->points.data(), cloud->size(), sizeof(PointTarget), detail::IntensityOffset<PointTarget>::get(), leaf,
                             q_last_curr_wxyz, t_last_curr, scan_period, (uint64_t)(uintptr_t)cloud.get());
  }
  // fast_gicp_impl.hpp:49-57.  The cloud handles follow the device-side swap, so that align() copies
  // the per-point fields of the cloud that is now the source (same point type), or none (different
  // point types / a filtered cloud: its centroids have no per-point fields).
  void swapSourceAndTarget() {
    detail::check(core_.ctx_, rgc_reg_swap_source_and_target(core_.reg_));
    std::swap(core_.n_src_, core_.n_tgt_);
    detail::swap_handles(source_, target_);
    std::swap(filtered_source_, filtered_target_);
  }
  void clearSource() { detail::check(core_.ctx_, rgc_reg_clear_source(core_.reg_)); source_.reset(); }
  void clearTarget() { detail::check(core_.ctx_, rgc_reg_clear_target(core_.reg_)); target_.reset(); }

  // covariances as column-major 4x4 doubles (Eigen::Matrix4d), 16 per point
  void setSourceCovariances(const std::vector<double>& m4x4) { detail::check(core_.ctx_, rgc_reg_set_source_covs(core_.reg_, m4x4.data(), m4x4.size() / 16)); }
  void setTargetCovariances(const std::vector<double>& m4x4) { detail::check(core_.ctx_, rgc_reg_set_target_covs(core_.reg_, m4x4.data(), m4x4.size() / 16)); }
  std::vector<double> getSourceCovariances() {
    std::vector<double> m(16 * core_.n_src_);
    detail::check(core_.ctx_, rgc_reg_get_source_covs(core_.reg_, m.data(), core_.n_src_));
    return m;
  }
  std::vector<double> getTargetCovariances() {
    std::vector<double> m(16 * core_.n_tgt_);
    detail::check(core_.ctx_, rgc_reg_get_target_covs(core_.reg_, m.data(), core_.n_tgt_));
    return m;
  }

  // pcl::Registration::align(output[, guess])
  void align(PointCloudSource& output) { align(output, identity4()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    std::vector<float> pts(4 * core_.n_src_);
    core_.align(guess.data(), final_.data(), pts.data());
    if (source_ && !filtered_source_) output = *source_;  // copies the non-geometric fields, like PCL
    else output = PointCloudSource();                     // filtered source: centroids have no per-point fields to copy
    output.resize(core_.n_src_);
    for (size_t i = 0; i < core_.n_src_; i++) {
      output[i].x = pts[4 * i];
      output[i].y = pts[4 * i + 1];
      output[i].z = pts[4 * i + 2];
      output[i].w = 1.f;
    }
  }
  const Matrix4& getFinalTransformation() const { return final_; }
  bool hasConverged() const { return core_.res_.converged != 0; }
  double getFitnessScore(double max_range = DBL_MAX) { return core_.fitness(max_range); }
  std::array<double, 36> getFinalHessian() const {
    std::array<double, 36> h;
    for (int i = 0; i < 36; i++) h[i] = core_.res_.final_hessian[i];
    return h;
  }
  // LsqRegistration::evaluateCost (lsq_registration_impl.hpp:48-51): pose cast to float first
  double evaluateCost(const Matrix4& relative_pose, double* H36 = nullptr, double* b6 = nullptr) {
    double T[16];
    for (int i = 0; i < 16; i++) T[i] = (double)relative_pose[i];
    return core_.linearize(T, H36, b6);
  }
  const rgc_result& lastResult() const { return core_.res_; }

 protected:
  detail::Core& core() { return core_; }

 private:
  detail::Core core_;
  PointCloudSourceConstPtr source_;
  bool filtered_source_ = false, filtered_target_ = false;
  PointCloudTargetConstPtr target_;
  Matrix4 final_ = identity4();
};

// fast_gicp::FastVGICP (fast_vgicp.hpp:24-80): voxelised correspondences, what RGC_odometer.cpp:998 uses
enum class NeighborSearchMethod { DIRECT27, DIRECT7, DIRECT1 };                       // gicp_settings.hpp:8
enum class VoxelAccumulationMode { ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE };     // gicp_settings.hpp:10
template <class PointSource, class PointTarget>
class FastVGICP : public FastGICP<PointSource, PointTarget> {
 public:
  explicit FastVGICP(int device = 0) : FastGICP<PointSource, PointTarget>(device) { push(); }
  void setResolution(double r) { res_ = r; push(); }
  void setNeighborSearchMethod(NeighborSearchMethod m) { search_ = (int)m; push(); }
  void setVoxelAccumulationMode(VoxelAccumulationMode m) { mode_ = (int)m; push(); }

 private:
  void push() { detail::check(this->core().ctx_, rgc_reg_set_vgicp(this->core().reg_, 1, res_, search_, mode_)); }
  double res_ = 1.0;
  int search_ = 2, mode_ = 0;
};

#else  // RGC_WITH_PCL ---------------------------------------------------------------------------
}  // namespace rgc
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/registration.h>
#include <Eigen/Core>
namespace rgc {

template <class PointSource, class PointTarget>
class FastGICP : public pcl::Registration<PointSource, PointTarget, float> {
 public:
  using Base = pcl::Registration<PointSource, PointTarget, float>;
  using Matrix4 = typename Base::Matrix4;
  using PointCloudSource = typename Base::PointCloudSource;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = typename Base::PointCloudTarget;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

  explicit FastGICP(int device = 0) : core_(device) {
    this->reg_name_ = "FastGICP";
    this->max_iterations_ = 64;                 // lsq_registration_impl.hpp:11
    this->transformation_epsilon_ = 5e-4;       // :13
    this->corr_dist_threshold_ = FLT_MAX;       // fast_gicp_impl.hpp:18
  }
  void setNumThreads(int) {}
  void setCorrespondenceRandomness(int k) { core_.prm_.k_correspondences = k; }
  void setRegularizationMethod(RegularizationMethod m) { core_.prm_.regularization = (int)m; }
  void setRotationEpsilon(double e) { core_.prm_.rotation_epsilon = e; }
  void setInitialLambdaFactor(double f) { core_.prm_.lm_init_lambda_factor = f; }
  void setDebugPrint(bool on) { core_.prm_.lm_debug_print = on ? 1 : 0; }
  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const { return final_hessian_; }

  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (this->input_ == cloud) return;  // fast_gicp_impl.hpp:73-75
    Base::setInputSource(cloud);
    core_.setSource(cloud->points.data(), cloud->size(), sizeof(PointSource), (uint64_t)(uintptr_t)cloud.get());
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (this->target_ == cloud) return;  // :84-86
    Base::setInputTarget(cloud);
    core_.setTarget(cloud->points.data(), cloud->size(), sizeof(PointTarget), (uint64_t)(uintptr_t)cloud.get());
  }
  double getFitnessScore(double max_range = DBL_MAX) { return core_.fitness(max_range); }

 protected:
  void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    core_.prm_.max_iterations = this->max_iterations_;
    core_.prm_.transformation_epsilon = this->transformation_epsilon_;
    core_.prm_.max_correspondence_distance = (float)this->corr_dist_threshold_;
    core_.push();
    std::vector<float> pts(4 * core_.n_src_);
    Eigen::Matrix4f g = guess, f;
    core_.align(g.data(), f.data(), pts.data());
    this->final_transformation_ = f;
    this->converged_ = core_.res_.converged != 0;
    this->nr_iterations_ = core_.res_.iterations;
    for (int i = 0; i < 36; i++) final_hessian_.data()[i] = core_.res_.final_hessian[i];
    for (size_t i = 0; i < core_.n_src_; i++) {
      output.points[i].x = pts[4 * i];
      output.points[i].y = pts[4 * i + 1];
      output.points[i].z = pts[4 * i + 2];
    }
  }

 private:
  detail::Core core_;
  Eigen::Matrix<double, 6, 6> final_hessian_ = Eigen::Matrix<double, 6, 6>::Identity();
};
#endif

}  // namespace rgc
