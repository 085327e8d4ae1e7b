// rgc_preprocess.cuh — the step in front of the registration, included by
// rgc_gicp.cu only: per-point motion compensation ("de-skew", RGC_odometer.cpp:1441-1481) and
// pcl::VoxelGrid centroid down-sampling (call sites RGC_odometer.cpp:975-991), so that a raw sweep
// goes raw -> de-skewed -> voxel-filtered -> Morton-sorted registration input without leaving HBM.
//
// Voxel filter on the device: voxel index per point (the reference's float arithmetic) -> stable
// LSD radix sort of (index, point id) -> ordered run heads by a block scan -> one thread per voxel
// adds its points in ascending input order, in float, and divides by float(count): the same order
// and arithmetic as the oracle's stable restatement, hence bit-identical centroids, emitted in
// ascending voxel-index order like PCL's.
#pragma once
#include "rgc_common.cuh"

namespace rgc {

constexpr size_t kNoIntensity = ~(size_t)0;

struct DeskewParams {
  int enabled;
  double iw, ix, iy, iz;  // q_last_curr^-1 (Eigen: conjugate / squaredNorm)
  double tx, ty, tz;      // t_last_curr
  float scan_period;
};

// Eigen 3.3 QuaternionBase::slerp(s, other) from the identity, then _transformVector on (p - s t);
// every double operation explicitly rounded (no FMA contraction), as the oracle is compiled.
RGC_HD void deskew_point(const DeskewParams& D, float inten, float& x, float& y, float& z) {
  const float sf = 1 - (inten - (float)(int)inten) / D.scan_period;  // float arithmetic (SCAN_PERIOD is a float)
  const double s = (double)sf;
  const double d = D.iw, absd = fabs(d);
  double scale0, scale1;
  if (absd >= 1.0 - 2.220446049250313e-16) {
    scale0 = dsub(1.0, s);
    scale1 = s;
  } else {
    const double theta = acos(absd);
    const double sin_theta = sin(theta);
    scale0 = sin(dmul(dsub(1.0, s), theta)) / sin_theta;
    scale1 = sin(dmul(s, theta)) / sin_theta;
  }
  if (d < 0.0) scale1 = -scale1;
  const double qw = dadd(scale0, dmul(scale1, D.iw)), qx = dmul(scale1, D.ix), qy = dmul(scale1, D.iy), qz = dmul(scale1, D.iz);
  const double vx = dsub((double)x, dmul(s, D.tx)), vy = dsub((double)y, dmul(s, D.ty)), vz = dsub((double)z, dmul(s, D.tz));
  double ux = dsub(dmul(qy, vz), dmul(qz, vy)), uy = dsub(dmul(qz, vx), dmul(qx, vz)), uz = dsub(dmul(qx, vy), dmul(qy, vx));
  ux = dadd(ux, ux);
  uy = dadd(uy, uy);
  uz = dadd(uz, uz);
  x = (float)dadd(dadd(vx, dmul(qw, ux)), dsub(dmul(qy, uz), dmul(qz, uy)));
  y = (float)dadd(dadd(vy, dmul(qw, uy)), dsub(dmul(qz, ux), dmul(qx, uz)));
  z = (float)dadd(dadd(vz, dmul(qw, uz)), dsub(dmul(qx, uy), dmul(qy, ux)));
}

#if defined(__CUDACC__)
// raw[n] (any PCL stride, xyz at byte 0, intensity at `inten_off` or absent) -> float4 (x, y, z, intensity),
// optionally de-skewed, plus per-block min/max partials of the OUTPUT coordinates
__global__ void __launch_bounds__(256) k_pre_ingest(const unsigned char* __restrict__ raw, size_t stride, size_t inten_off, int n, DeskewParams D,
                                                    float4* __restrict__ out, float* __restrict__ bbox_partials) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned char* rec = raw + (size_t)i * stride;
    const float* p = reinterpret_cast<const float*>(rec);
    float x = p[0], y = p[1], z = p[2];
    const float inten = inten_off == kNoIntensity ? 0.f : *reinterpret_cast<const float*>(rec + inten_off);
    if (D.enabled) deskew_point(D, inten, x, y, z);
    out[i] = make_float4(x, y, z, inten);
    bad |= !(isfinite(x) && isfinite(y) && isfinite(z));
    mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
    mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
    mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
  }
  bbox_block_reduce(mn, mx, bad, bbox_partials);
}

#endif  // __CUDACC__

struct VgGeom {  // pcl::VoxelGrid::applyFilter: min_b_, divb_mul_, inverse_leaf_size_
  float inv_leaf;
  int min_b[3];
  int mul1, mul2;
};

// pcl::VoxelGrid voxel index of a point (float arithmetic of filters/impl/voxel_grid.hpp)
RGC_HD unsigned vg_index(const VgGeom& g, float x, float y, float z) {
  const int i0 = (int)fsub(floorf(fmul(x, g.inv_leaf)), (float)g.min_b[0]);
  const int i1 = (int)fsub(floorf(fmul(y, g.inv_leaf)), (float)g.min_b[1]);
  const int i2 = (int)fsub(floorf(fmul(z, g.inv_leaf)), (float)g.min_b[2]);
  return (unsigned)(i0 + i1 * g.mul1 + i2 * g.mul2);
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) k_vg_keys(const float4* __restrict__ pts, int n, VgGeom g, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  keys[i] = (uint64_t)vg_index(g, p.x, p.y, p.z);
  vals[i] = (uint32_t)i;
}

// ---- ordered run heads: per-block head counts, a one-block scan of the counts, ordered write ----
constexpr int kScanItems = 4;  // elements per thread, 256 threads -> 1024 per block
__device__ __forceinline__ bool vg_is_head(const uint64_t* __restrict__ keys, int i, int n) { return i < n && (i == 0 || keys[i] != keys[i - 1]); }

__global__ void __launch_bounds__(256) k_vg_head_count(const uint64_t* __restrict__ keys, int n, unsigned int* __restrict__ block_counts) {
  const int base = blockIdx.x * 256 * kScanItems;
  int c = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; u++) c += vg_is_head(keys, base + u * 256 + threadIdx.x, n) ? 1 : 0;
  __shared__ int sm[8];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; w++) t += sm[w];
    block_counts[blockIdx.x] = (unsigned)t;
  }
}
// exclusive scan of block_counts in place (one block; nblk <= a few 10^4), total -> block_counts[nblk]
__global__ void __launch_bounds__(1024) k_vg_scan_blocks(unsigned int* __restrict__ block_counts, int nblk) {
  __shared__ unsigned int warp_tot[32];
  __shared__ unsigned int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nblk; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const unsigned v = i < nblk ? block_counts[i] : 0u;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_tot[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;  // inclusive
    }
    __syncthreads();
    const unsigned before = carry + (warp > 0 ? warp_tot[warp - 1] : 0u) + (x - v);
    if (i < nblk) block_counts[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_counts[nblk] = carry;
}
__global__ void __launch_bounds__(256) k_vg_heads(const uint64_t* __restrict__ keys, int n, const unsigned int* __restrict__ block_offsets, int* __restrict__ heads) {
  // elements are visited in index order: item u of thread t is element base + u * 256 + t, so ranks are
  // computed per pass of 256 consecutive elements
  const int base = blockIdx.x * 256 * kScanItems;
  __shared__ int warp_cnt[8];
  __shared__ int pass_base;
  if (threadIdx.x == 0) pass_base = (int)block_offsets[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = 0; u < kScanItems; u++) {
    const int i = base + u * 256 + threadIdx.x;
    const bool h = vg_is_head(keys, i, n);
    const unsigned bal = __ballot_sync(0xffffffffu, h);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = pass_base;
    for (int w = 0; w < warp; w++) before += warp_cnt[w];
    if (h) heads[before + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; w++) t += warp_cnt[w];
      pass_base += t;
    }
    __syncthreads();
  }
}

// pcl::CentroidPoint<PointXYZI>: float sums of x, y, z, intensity over the voxel's points in ascending
// input order (vals are ascending inside a run: the sort is stable), divided by float(count)
__global__ void __launch_bounds__(128) k_vg_centroid(const float4* __restrict__ pts, const uint32_t* __restrict__ vals, const int* __restrict__ heads, int nv, int n,
                                                     float4* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int s = heads[v], e = v + 1 < nv ? heads[v + 1] : n;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  // gathers in batches of 8 (independent loads in flight), additions strictly in input order
  for (int j0 = s; j0 < e; j0 += 8) {
    float4 p[8];
#pragma unroll
    for (int u = 0; u < 8; u++) p[u] = j0 + u < e ? __ldg(&pts[vals[j0 + u]]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (j0 + u < e) {
        sx = fadd(sx, p[u].x);
        sy = fadd(sy, p[u].y);
        sz = fadd(sz, p[u].z);
        si = fadd(si, p[u].w);
      }
  }
  const float cnt = (float)(e - s);
  out[v] = make_float4(sx / cnt, sy / cnt, sz / cnt, si / cnt);
}

#endif  // __CUDACC__

}  // namespace rgc
