// rgc_common.cuh — shared definitions for the B200 scan-matching kernels.
//
// Everything marked RGC_HD is plain arithmetic that compiles both for sm_100a (nvcc) and for
// the host (g++).  The host build exists ONLY for tests/hostsim (a CPU simulation of the
// kernels' per-thread logic, so the search / algebra can be checked against the oracle in a
// container without a GPU).  The product library never runs these functions on the CPU.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define RGC_HD __host__ __device__ __forceinline__
#define RGC_D __device__ __forceinline__
#define RGC_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define RGC_HD inline
#define RGC_D inline
#define RGC_HD_NOINLINE inline
#endif

namespace rgc {

// ---- float arithmetic with pinned rounding (no FMA contraction) -------------------------------
// The reference's float expressions (flann::L2_Simple distance, Isometry3f * Vector4f,
// scanRegistration.cpp curvature sums) are evaluated on x86-64 without FMA; kNN indices and
// feature labels must be bit-exact, so every float product/sum on those paths goes through
// these helpers (round-to-nearest, never fused).
RGC_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
RGC_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
RGC_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
RGC_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
RGC_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
RGC_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}

// fused multiply-add, always: the fp64 algebra of linearize / compute_error is written with explicit
// dfma / dmul / dadd so that every kernel that inlines it (single registration, batched, voxelised)
// performs the SAME roundings — left to the compiler, the choice of which product of a sum gets fused
// varies with the inlining context and two kernels disagree in the last bit.
RGC_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

RGC_HD int f2i_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int i;
  std::memcpy(&i, &f, 4);
  return i;
#endif
}
RGC_HD float i2f_bits(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f;
  std::memcpy(&f, &i, 4);
  return f;
#endif
}

// float -> double (exact).  Widening by hand with integer operations instead of the quarter-rate conversion
// instruction (60 per point in k_covariance) was measured and is SLOWER (0.46 -> 0.55 ms at 8 M points): the
// kernels wait on memory, not on the conversion pipe (profiles/README.md).
RGC_HD double f2d(float f) { return (double)f; }

struct F4 {  // layout-compatible with float4
  float x, y, z, w;
};

// squared distance exactly as flann::L2_Simple<float> accumulates it: ((dx*dx)+dy*dy)+dz*dz
RGC_HD float dist2_ref(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = fsub(qx, px), dy = fsub(qy, py), dz = fsub(qz, pz);
  float d = fmul(dx, dx);
  d = fadd(d, fmul(dy, dy));
  d = fadd(d, fmul(dz, dz));
  return d;
}

// float Isometry3f * Vector4f with w = 1 (fast_gicp_impl.hpp:131): ((r0*x + r1*y) + r2*z) + t
// T: row-major 3x4 floats.
RGC_HD void transform_f(const float* T, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = fadd(fadd(fadd(fmul(T[0], x), fmul(T[1], y)), fmul(T[2], z)), T[3]);
  oy = fadd(fadd(fadd(fmul(T[4], x), fmul(T[5], y)), fmul(T[6], z)), T[7]);
  oz = fadd(fadd(fadd(fmul(T[8], x), fmul(T[9], y)), fmul(T[10], z)), T[11]);
}

#if defined(__CUDACC__)
// Programmatic dependent launch (sm_90+ `griddepcontrol`).  The hot chains of this library are strings of short
// dependent kernels (a cloud build: keys -> five sort passes -> tables; an LM step: correspondences -> on-demand
// k-NN -> covariances -> linearize), each boundary costing a drain + launch gap of a few microseconds.  A kernel
// launched with cudaLaunchAttributeProgrammaticStreamSerialization (launch_pdl, rgc_gicp.cu) may be scheduled
// while its predecessor is still running; pdl_enter() is therefore the FIRST statement of every kernel on those
// chains, before any early return: `wait` blocks until the predecessor grid has completed and its writes are
// visible (so the kernel body sees exactly what it would see in plain stream order), `launch_dependents` lets the
// successor's blocks become resident behind this grid's last wave.  Without the launch attribute both
// instructions are no-ops, so the same kernels can be launched with <<<>>> elsewhere.
__device__ __forceinline__ void pdl_enter() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

enum RegMethod { REG_NONE = 0, REG_MIN_EIG = 1, REG_NORMALIZED_MIN_EIG = 2, REG_PLANE = 3, REG_FROBENIUS = 4 };

}  // namespace rgc
