// rgc_features.cu — A-LOAM scan-line feature extraction + ground-plane fit on B200 (sm_100a).
//
// Replaces the numeric body of ScanRegistration::laserCloudHandler
// (/root/reference/rgc_slam/src/scanRegistration.cpp:110-663, :732-763).  One scan is ~29k points
// (~1 MB): far too small for 148 SMs, so the unit of parallelism is the BATCH — one CUDA block
// per scan for the order-dependent steps (ring bucketing, ground fit, list compaction), one block
// per (scan, ring) for the sort + greedy selection, one thread per point for the stencils.
//
// Labels must be bit-exact, so every float/double expression below reproduces the reference's
// evaluation order and float/double promotions with explicitly rounded operations (no FMA
// contraction) — see rgc_common.cuh and SURVEY.md Appendix A.9-A.15.
//
//   k_feat_rings      :110-230  range gate, ring id, relTime, stable bucket by ring
//   k_feat_pass_a     :233-268  range, incidence angle (r < 2 m), int-truncated intensity smoothing
//   k_feat_pass_b     :270-306  cloudCurvature, intensityCurvature, cloudCurvature2, weights
//   k_feat_occlusion  :433-456  occlusion / parallel-beam masking
//   k_feat_ground     :307-431  ground marking + weighted PCA plane (11 doubles)
//   k_feat_select     :469-644  per ring: 6 sextants in order, sort + greedy picks
//   k_feat_compact    :645-663  feature lists in push_back order, intensity merge flag
#include <cfloat>
#include <cstdio>
#include <vector>

#include "../../include/rgc_features.h"
#include "rgc_common.cuh"
#include "rgc_ctx.hpp"
#include "rgc_math.cuh"

namespace rgc {
namespace {

constexpr int kMaxRings = 64;
constexpr int kSegCap = 2048;  // longest sextant the selection kernel sorts in shared memory
constexpr double kPi = 3.14159265358979323846;

struct FeatArrays {  // device pointers, per-point arrays in the "8 slots of slack per scan" layout
  float4* cloud;
  int* src_index;
  int* intensity2;  // raw integer intensity, ring order (intensity_num2)
  int* intensity_num;
  float *range_vec, *scan_angle, *curvature, *inten_curvature, *curvature2, *distance_source, *other_source;
  int *label, *inten_label, *neighbor_picked, *inten_neighbor_picked, *ground_marked;
  // per scan
  int *cloud_size, *scan_start, *scan_end, *ground_size, *inten_merged;
  double* groundparam;
  // per (scan, ring, sextant) pick buffers
  int *seg_sharp, *seg_less, *seg_flat, *seg_inten, *seg_inten_less;  // [seg][20|22|40|20|21]
  int* seg_counts;                                                   // [seg][5]
  // compacted lists
  int *corner_sharp, *corner_less_sharp, *surf_flat, *inten_sharp, *inten_less_sharp;
  float *corner_sharp_w, *surf_flat_w, *inten_sharp_w;
  int* list_counts;  // [scan][5]
  // surfPointsLessFlatScan (:586-592) in the per-point capacity layout, GroundPoints (:338) with `ground_cap` slots per scan
  int *surf_less_flat, *n_surf_less_flat, *ground_points;
  int ground_cap;
};

__device__ __forceinline__ float absf(float v) { return v < 0 ? -v : v; }

// ------------------------------------------------------------------------------------------------
// One block per scan.  Stable partition of the kept points by ring id (scanRegistration.cpp:135-230).
__global__ void __launch_bounds__(256) k_feat_rings(const float4* __restrict__ raw, const int* __restrict__ scan_offsets, int n_rings, float th1, float th2,
                                                    int* __restrict__ tmp_ring, float* __restrict__ tmp_inten, int* __restrict__ max_seg, FeatArrays A) {
  const int b = blockIdx.x;
  const int in0 = scan_offsets[b], n = scan_offsets[b + 1] - in0;
  const int out0 = in0 + 8 * b;
  const float4* P = raw + in0;
  int* ring = tmp_ring + in0;
  float* inten = tmp_inten + in0;
  __shared__ int s_first, s_last, s_half;
  __shared__ int s_ring_cnt[kMaxRings], s_ring_off[kMaxRings + 1], s_ring_run[kMaxRings];
  __shared__ int s_warp_cnt[8][kMaxRings];
  __shared__ float s_start_ori, s_end_ori;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_first = 0x7fffffff;
    s_last = -1;
    s_half = 0x7fffffff;
  }
  if (tid < kMaxRings) s_ring_cnt[tid] = 0;
  __syncthreads();
  // ---- removeClosedPointCloud (:732-763) + NaN removal (:112): keep flag, first / last kept
  const float th1sq = fmul(th1, th1), th2sq = fmul(th2, th2);
  for (int i = tid; i < n; i += blockDim.x) {
    const float4 p = P[i];
    bool keep = isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
    if (keep) {
      const float dis = fadd(fadd(fmul(p.x, p.x), fmul(p.y, p.y)), fmul(p.z, p.z));
      if (dis < th1sq || dis > th2sq) keep = false;
      if (p.x < 0 && (double)absf(p.y) < 0.5) keep = false;
    }
    ring[i] = keep ? 0 : -2;
    if (keep) {
      atomicMin(&s_first, i);
      atomicMax(&s_last, i);
    }
  }
  __syncthreads();
  if (s_last < 0) {  // nothing survives the range gate
    if (tid == 0) A.cloud_size[b] = 0;
    if (tid < kMaxRings) {
      A.scan_start[b * kMaxRings + tid] = 5;
      A.scan_end[b * kMaxRings + tid] = -5;
    }
    return;
  }
  if (tid == 0) {  // :117-127
    const float4 f = P[s_first], l = P[s_last];
    float startOri = -(float)atan2((double)f.y, (double)f.x);
    float endOri = (float)(-(double)(float)atan2((double)l.y, (double)l.x) + 2 * kPi);
    if ((double)fsub(endOri, startOri) > 3 * kPi)
      endOri = (float)((double)endOri - 2 * kPi);
    else if ((double)fsub(endOri, startOri) < kPi)
      endOri = (float)((double)endOri + 2 * kPi);
    s_start_ori = startOri;
    s_end_ori = endOri;
  }
  __syncthreads();
  const float startOri = s_start_ori, endOri = s_end_ori;
  // ---- ring id (:142-183) and the first index that flips halfPassed (:187-196)
  for (int i = tid; i < n; i += blockDim.x) {
    if (ring[i] < 0) continue;
    const float4 p = P[i];
    const float hyp = __fsqrt_rn(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));
    const float va = (float)((double)fmul((float)atan((double)__fdiv_rn(p.z, hyp)), 180.f) / kPi);
    int id = 0;
    bool ok = true;
    if (n_rings == 16) {
      id = (int)((double)fmul(fadd(va, 15.f), 0.5f) + 0.5);
      ok = !(id > 15 || id < 0);
    } else if (n_rings == 32) {
      id = (int)(((double)va + 92.0 / 3.0) * 3.0 / 4.0);
      ok = !(id > 31 || id < 0);
    } else {
      if ((double)va >= -8.83)
        id = (int)((2 - (double)va) * 3.0 + 0.5);
      else
        id = 32 + (int)((-8.83 - (double)va) * 2.0 + 0.5);
      ok = !((double)va > 2 || (double)va < -24.33 || id > 50 || id < 0);
    }
    ring[i] = ok ? id : -1;
    if (ok) {
      atomicAdd(&s_ring_cnt[id], 1);
      float ori = -(float)atan2((double)p.y, (double)p.x);
      if ((double)ori < (double)startOri - kPi / 2)
        ori = (float)((double)ori + 2 * kPi);
      else if ((double)ori > (double)startOri + kPi * 3 / 2)
        ori = (float)((double)ori - 2 * kPi);
      if ((double)fsub(ori, startOri) > kPi) atomicMin(&s_half, i);
    }
  }
  __syncthreads();
  const int half_at = s_half;  // points with index <= half_at take the !halfPassed branch
  // ---- relTime -> intensity = scanID + 0.1 * relTime (:186-210)
  for (int i = tid; i < n; i += blockDim.x) {
    const int id = ring[i];
    if (id < 0) continue;
    const float4 p = P[i];
    float ori = -(float)atan2((double)p.y, (double)p.x);
    if (i <= half_at) {
      if ((double)ori < (double)startOri - kPi / 2)
        ori = (float)((double)ori + 2 * kPi);
      else if ((double)ori > (double)startOri + kPi * 3 / 2)
        ori = (float)((double)ori - 2 * kPi);
    } else {
      ori = (float)((double)ori + 2 * kPi);
      if ((double)ori < (double)endOri - kPi * 3 / 2)
        ori = (float)((double)ori + 2 * kPi);
      else if ((double)ori > (double)endOri + kPi / 2)
        ori = (float)((double)ori - 2 * kPi);
    }
    const float relTime = __fdiv_rn(fsub(ori, startOri), fsub(endOri, startOri));
    inten[i] = (float)dadd((double)id, dmul(0.1, (double)relTime));  // scanID + scanPeriod * relTime (:210)
  }
  // ---- ring offsets (:221-230)
  if (tid == 0) {
    int acc = 0;
    for (int r = 0; r < n_rings; r++) {
      s_ring_off[r] = acc;
      acc += s_ring_cnt[r];
    }
    s_ring_off[n_rings] = acc;
    A.cloud_size[b] = acc;
  }
  if (tid < kMaxRings) s_ring_run[tid] = 0;
  __syncthreads();
  if (tid < kMaxRings) {
    const bool in = tid < n_rings;
    A.scan_start[b * kMaxRings + tid] = in ? s_ring_off[tid] + 5 : 0;
    A.scan_end[b * kMaxRings + tid] = in ? s_ring_off[tid + 1] - 5 : 0;
    if (in) {  // longest sextant of the batch: sizes the selection kernel's shared-memory sort
      const int span = s_ring_off[tid + 1] - s_ring_off[tid] - 10;
      if (span >= 10) atomicMax(max_seg, span / 6 + 2);
    }
  }
  // ---- stable scatter, chunk by chunk in firing order
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < n; base += blockDim.x) {
    for (int j = tid; j < 8 * kMaxRings; j += blockDim.x) (&s_warp_cnt[0][0])[j] = 0;
    __syncthreads();
    const int i = base + tid;
    const int id = i < n ? ring[i] : -1;
    const unsigned key = id >= 0 ? (unsigned)id : (1000u + lane);
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(peers & lt_mask);
    if (id >= 0 && rank == 0) s_warp_cnt[warp][id] = __popc(peers);
    __syncthreads();
    int dest = -1;
    if (id >= 0) {
      int before = s_ring_run[id];
      for (int w = 0; w < warp; w++) before += s_warp_cnt[w][id];
      dest = s_ring_off[id] + before + rank;
    }
    __syncthreads();
    if (tid < n_rings) {
      int add = 0;
      for (int w = 0; w < 8; w++) add += s_warp_cnt[w][tid];
      s_ring_run[tid] += add;
    }
    if (dest >= 0) {
      const float4 p = P[i];
      A.cloud[out0 + dest] = make_float4(p.x, p.y, p.z, inten[i]);
      A.src_index[out0 + dest] = i;
      A.intensity2[out0 + dest] = (int)p.w;  // point_intensity = laserCloudIn.points[i].intensity (:140)
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// thread per ordered point: :234-268
__global__ void __launch_bounds__(256) k_feat_pass_a(const int* __restrict__ scan_offsets, int n_scans, int total_out, FeatArrays A) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_out) return;
  // locate the scan (few scans per block: linear probe from an estimate is fine; use binary search)
  int lo = 0, hi = n_scans - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (scan_offsets[mid] + 8 * mid <= g) lo = mid; else hi = mid - 1;
  }
  const int b = lo, out0 = scan_offsets[b] + 8 * b, i = g - out0, size = A.cloud_size[b];
  if (i >= size) return;
  const float4* C = A.cloud + out0;
  const float4 p = C[i];
  const float range = __fsqrt_rn(fadd(fadd(fmul(p.x, p.x), fmul(p.y, p.y)), fmul(p.z, p.z)));
  A.range_vec[g] = range;
  float sa = 0.f;
  int inum = A.intensity2[g];
  if (i >= 5 && i < size - 5) {
    if (range < 2) {  // :241-254, double arithmetic, sequential sums
      const float4 a4 = C[i + 5], b4 = C[i - 5];
      const double a[3] = {a4.x, a4.y, a4.z}, bb[3] = {b4.x, b4.y, b4.z}, now[3] = {p.x, p.y, p.z};
      double c[3], ab[3], nc[3];
#pragma unroll
      for (int d = 0; d < 3; d++) {
        c[d] = dadd(a[d], bb[d]) / 2;
        ab[d] = dsub(a[d], bb[d]);
        nc[d] = dsub(now[d], c[d]);
      }
      const double nrm[3] = {dsub(dmul(ab[1], nc[2]), dmul(ab[2], nc[1])), dsub(dmul(ab[2], nc[0]), dmul(ab[0], nc[2])),
                             dsub(dmul(ab[0], nc[1]), dmul(ab[1], nc[0]))};
      const double dot = dadd(dadd(dmul(nrm[0], now[0]), dmul(nrm[1], now[1])), dmul(nrm[2], now[2]));
      const double n1 = sqrt(dadd(dadd(dmul(nrm[0], nrm[0]), dmul(nrm[1], nrm[1])), dmul(nrm[2], nrm[2])));
      const double n2 = sqrt(dadd(dadd(dmul(now[0], now[0]), dmul(now[1], now[1])), dmul(now[2], now[2])));
      sa = (float)(dot / dmul(n1, n2));
      if (sa < 0) sa = -sa;
    }
    if ((double)sa < 0.07 && range < 2) {  // :259-267, int truncation after every +=
      const int* I2 = A.intensity2 + out0;
      inum = (int)dmul(0.9, (double)I2[i]);
      for (int j = -5; j < 6; j++)
        if (j != 0) inum = (int)dadd((double)inum, dmul(0.005, (double)I2[i + j]));
    }
  }
  A.scan_angle[g] = sa;
  A.intensity_num[g] = inum;
}

// thread per ordered point: :270-306
__global__ void __launch_bounds__(256) k_feat_pass_b(const int* __restrict__ scan_offsets, int n_scans, int total_out, FeatArrays A) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_out) return;
  int lo = 0, hi = n_scans - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (scan_offsets[mid] + 8 * mid <= g) lo = mid; else hi = mid - 1;
  }
  const int b = lo, out0 = scan_offsets[b] + 8 * b, i = g - out0, size = A.cloud_size[b];
  if (i < 5 || i >= size - 5) return;
  const float4* C = A.cloud + out0 + i;
  const int* q = A.intensity_num + out0 + i;
  const float* r = A.range_vec + out0 + i;
  float dx, dy, dz;
  {
    float4 c0 = C[-5], c1 = C[-4], c2 = C[-3], c3 = C[-2], c4 = C[-1], c5 = C[0], c6 = C[1], c7 = C[2], c8 = C[3], c9 = C[4], c10 = C[5];
#define RGC_DIFF(f) fadd(fadd(fadd(fadd(fadd(fsub(fadd(fadd(fadd(fadd(c0.f, c1.f), c2.f), c3.f), c4.f), fmul(10.f, c5.f)), c6.f), c7.f), c8.f), c9.f), c10.f)
    dx = RGC_DIFF(x);
    dy = RGC_DIFF(y);
    dz = RGC_DIFF(z);
#undef RGC_DIFF
  }
  const float diffI = (float)(q[-5] + q[-4] + q[-3] + q[-2] + q[-1] - 10 * q[0] + q[1] + q[2] + q[3] + q[4] + q[5]);
  const float range = r[0];
  float dis_factor = (float)(2.0 / (1.0 + (double)range / 20.0));
  if ((double)dis_factor < 0.2) dis_factor = 0.2f;
  A.curvature[g] = fmul(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)), dis_factor);
  A.distance_source[g] = (float)(0.5 + (double)dis_factor);
  const float sa = A.scan_angle[g];
  float inten_factor;
  if ((double)sa < 0.07 && range < 2) {
    inten_factor = (float)((double)fmul(sa, 10.f) + 0.6);
    A.inten_curvature[g] = (float)dmul(dadd((double)sa, 0.3), (double)diffI);
  } else {
    inten_factor = 3;
    A.inten_curvature[g] = diffI;
  }
  A.other_source[g] = inten_factor;
  const float s5 = fadd(fadd(fadd(fadd(r[-5], r[-4]), r[-3]), r[-2]), r[-1]);
  double dr = dsub((double)s5, dmul(10.0, (double)r[0]));
  dr = dadd(dr, (double)r[1]);
  dr = dadd(dr, (double)r[2]);
  dr = dadd(dr, (double)r[3]);
  dr = dadd(dr, (double)r[4]);
  dr = dadd(dr, (double)r[5]);
  A.curvature2[g] = absf(fmul((float)dr, dis_factor));
}

// thread per ordered point: :433-456 (idempotent stores, order independent)
__global__ void __launch_bounds__(256) k_feat_occlusion(const int* __restrict__ scan_offsets, int n_scans, int total_out, FeatArrays A) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_out) return;
  int lo = 0, hi = n_scans - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (scan_offsets[mid] + 8 * mid <= g) lo = mid; else hi = mid - 1;
  }
  const int b = lo, out0 = scan_offsets[b] + 8 * b, i = g - out0, size = A.cloud_size[b];
  if (i < 5 || i >= size - 5) return;
  const float d1 = A.range_vec[g], d2 = A.range_vec[g + 1];
  int* np = A.neighbor_picked + g;
  if ((double)fsub(d1, d2) > dmul(0.04, (double)d2)) {
    for (int l = -5; l <= 0; l++) np[l] = 1;
  } else if ((double)fsub(d2, d1) > dmul(0.04, (double)d1)) {
    for (int l = 1; l <= 6; l++) np[l] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// block-wide deterministic sum of NV doubles (fixed tree), result broadcast through smem
template <int NV>
__device__ void block_sum(double* v, double* out /*smem[NV]*/) {
  __shared__ double sm[8][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double x = v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp][j] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sm[w][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// One block per scan: ground marking + weighted plane (:307-431).  The reference appends samples
// sequentially (duplicates included) and sums them in that order; here every (ring, column, n)
// candidate is evaluated independently and the sums are fixed-order tree reductions: flags are
// identical, the 11 doubles agree to rounding (they are not labels).
__global__ void __launch_bounds__(256) k_feat_ground(const int* __restrict__ scan_offsets, int n_rings, FeatArrays A) {
  const int b = blockIdx.x, out0 = scan_offsets[b] + 8 * b;
  const float Ground_scan_range[8] = {2.66f, 3.04f, 3.56f, 4.30f, 5.44f, 7.41f, 11.63f, 27.12f};
  const int groundScanInd = 7;
  const double laderH = 0.56;
  const float4* C = A.cloud + out0;
  const float* R = A.range_vec + out0;
  int* GM = A.ground_marked + out0;
  __shared__ double s_out[12];
  __shared__ double s_center[3], s_n[3];
  const int* sst = A.scan_start + b * kMaxRings;
  const int* sen = A.scan_end + b * kMaxRings;

  // a sample = (ring i, column col, offset n) passing both tests; enumerate them identically in
  // each of the three passes
  auto for_each_sample = [&](auto&& fn) {
    for (int i = 0; i < groundScanInd && i < n_rings; i++) {
      const int ring0 = sst[i] - 5, ring_n = sen[i] + 5 - ring0;
      if (ring_n < 5) continue;  // reference: size_t wrap; generators guarantee >= 11
      const float th = (float)(0.8 * (1.0 + i / (groundScanInd - 1)));
      const double w = 1.5 - i / (groundScanInd - 1);
      for (int col = 5 + threadIdx.x; col < ring_n - 5; col += blockDim.x) {
        const int ci = ring0 + col;
        const float rc = R[ci];
        if (!(absf(fsub(rc, Ground_scan_range[i])) < th)) continue;
        if (!((double)C[ci].z < 0.3)) continue;
        fn(ci, -100, w);  // the column itself is marked (:332) but not appended
        for (int n = -5; n < 5; n++)
          if (absf(fsub(R[ci + n], rc)) < fmul(th, 0.5f)) fn(ci + n, n, w);
      }
    }
  };
  // pass 1: marks, weighted centre, counts
  double acc[5] = {0, 0, 0, 0, 0};
  for_each_sample([&](int idx, int n, double w) {
    GM[idx] = 1;
    if (n == -100) return;
    const float4 p = C[idx];
    acc[0] += w * (double)p.x;
    acc[1] += w * (double)p.y;
    acc[2] += w * (double)p.z;
    acc[3] += w;
    acc[4] += 1.0;
  });
  block_sum<5>(acc, s_out);
  const double groundweights = s_out[3];
  const int groundsize = (int)s_out[4];
  if (threadIdx.x == 0) A.ground_size[b] = groundsize;
  if (threadIdx.x < 3) s_center[threadIdx.x] = groundsize ? s_out[threadIdx.x] / groundweights : 0.0;
  __syncthreads();
  if (groundsize == 0) {
    if (threadIdx.x < 11) A.groundparam[b * 11 + threadIdx.x] = 0.0;
    return;
  }
  const double cx = s_center[0], cy = s_center[1], cz = s_center[2];
  // pass 2: weighted covariance about the centre
  double cov[6] = {0, 0, 0, 0, 0, 0};
  for_each_sample([&](int idx, int n, double w) {
    if (n == -100) return;
    const float4 p = C[idx];
    const double tx = (double)p.x - cx, ty = (double)p.y - cy, tz = (double)p.z - cz;
    cov[0] += w * tx * tx;
    cov[1] += w * tx * ty;
    cov[2] += w * tx * tz;
    cov[3] += w * ty * ty;
    cov[4] += w * ty * tz;
    cov[5] += w * tz * tz;
  });
  __syncthreads();
  block_sum<6>(cov, s_out);
  __shared__ double s_V[3][3];
  __shared__ int s_order[3];
  if (threadIdx.x == 0) {
    Sym3 Cm = {s_out[0] / groundweights, s_out[1] / groundweights, s_out[2] / groundweights,
               s_out[3] / groundweights, s_out[4] / groundweights, s_out[5] / groundweights};
    double w[3], V[3][3];
    eig_sym3(Cm, w, V);
    int o[3] = {0, 1, 2};  // ascending eigenvalues (SelfAdjointEigenSolver order)
    if (w[o[1]] < w[o[0]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    if (w[o[2]] < w[o[1]]) { int t = o[1]; o[1] = o[2]; o[2] = t; }
    if (w[o[1]] < w[o[0]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    for (int c = 0; c < 3; c++) {
      s_order[c] = o[c];
      for (int rr = 0; rr < 3; rr++) s_V[rr][c] = V[rr][c];
    }
    double n0 = V[0][o[0]], n1 = V[1][o[0]], n2 = V[2][o[0]];
    const double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    n0 /= nn; n1 /= nn; n2 /= nn;
    if (cx * n0 + cy * n1 + cz * n2 < 0) { n0 = -n0; n1 = -n1; n2 = -n2; }
    s_n[0] = n0; s_n[1] = n1; s_n[2] = n2;
  }
  __syncthreads();
  const double n0 = s_n[0], n1 = s_n[1], n2 = s_n[2];
  // pass 3: robust distance (:386-413)
  double ds[2] = {0, 0};
  for_each_sample([&](int idx, int n, double w) {
    if (n == -100) return;
    const float4 p = C[idx];
    double tx = (double)p.x - cx, ty = (double)p.y - cy, tz = (double)p.z - cz;
    const double tn = sqrt(tx * tx + ty * ty + tz * tz);
    if (tn > 0) { tx /= tn; ty /= tn; tz /= tn; }
    double dw = 1 - 100 * fabs(n0 * tx + n1 * ty + n2 * tz);
    if (dw < 0) dw = 0.1;
    ds[0] += dw;
    ds[1] += dw * (n0 * (double)p.x + n1 * (double)p.y + n2 * (double)p.z);
  });
  __syncthreads();
  block_sum<2>(ds, s_out);
  if (threadIdx.x == 0) {
    double groundsource1 = s_out[0];
    double distance = s_out[1] / groundsource1;
    groundsource1 = groundsource1 / groundsize;
    if ((distance / laderH) > 1.1 || (distance / laderH) < 0.9) distance = laderH;
    if (groundsource1 < 0.9) distance = 0.9 * laderH + 0.1 * distance;
    double* gp = A.groundparam + b * 11;
    gp[0] = n0; gp[1] = n1; gp[2] = n2;
    const int o1 = s_order[1], o2 = s_order[2];
    gp[3] = s_V[0][o1]; gp[4] = s_V[1][o1]; gp[5] = s_V[2][o1];
    gp[6] = s_V[0][o2]; gp[7] = s_V[1][o2]; gp[8] = s_V[2][o2];
    gp[9] = distance;
    gp[10] = 1 - groundsource1;
  }
}

// ------------------------------------------------------------------------------------------------
// bitonic sort of (key, index) ascending by (key, index) in shared memory, n padded to a power of two
__device__ void bitonic_sort(float* key, int* idx, int npad) {
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < npad; t += blockDim.x) {
        const int x = t ^ j;
        if (x > t) {
          const bool up = (t & k) == 0;
          const float a = key[t], b = key[x];
          const int ia = idx[t], ib = idx[x];
          const bool a_after_b = a > b || (a == b && ia > ib);
          if (a_after_b == up) {
            key[t] = b; key[x] = a;
            idx[t] = ib; idx[x] = ia;
          }
        }
      }
      __syncthreads();
    }
}

// One block per (scan, ring); the six sextants run in order because a pick suppresses up to five
// neighbours that may lie in the next sextant (:517-534).  Each sextant: its window [sp-5, ep+5] of
// points, curvatures, intensities and flags is staged in shared memory, two bitonic sorts (by
// cloudCurvature and by intensityCurvature, ties by index) run on all threads, then thread 0
// replays the greedy loops (:485-641) entirely out of shared memory and the flags are written back.
// (The first version ran the greedy loops against global memory: 87 % of the feature path's time.)
constexpr int kSelThreads = 64;
struct SelSmem {  // carve-up of the dynamic shared memory for a given sort capacity `cap`
  float* k1; int* i1; float* k2; int* i2;          // [cap] sort keys / indices
  float4* pt;                                      // [cap + 10] window points
  float *cv, *c2, *ci; int* in;                    // [cap + 10]
  signed char *np, *inp, *lb, *ilb, *gm;           // [cap + 10]
  __device__ SelSmem(unsigned char* base, int cap) {
    const int w = cap + 10;
    pt = reinterpret_cast<float4*>(base);
    k1 = reinterpret_cast<float*>(pt + w);
    i1 = reinterpret_cast<int*>(k1 + cap);
    k2 = reinterpret_cast<float*>(i1 + cap);
    i2 = reinterpret_cast<int*>(k2 + cap);
    cv = reinterpret_cast<float*>(i2 + cap);
    c2 = cv + w;
    ci = c2 + w;
    in = reinterpret_cast<int*>(ci + w);
    np = reinterpret_cast<signed char*>(in + w);
    inp = np + w;
    lb = inp + w;
    ilb = lb + w;
    gm = ilb + w;
  }
};
static size_t sel_smem_bytes(int cap) { return (size_t)(cap + 10) * (16 + 16 + 5) + (size_t)cap * 16 + 16; }

__global__ void __launch_bounds__(kSelThreads) k_feat_select(const int* __restrict__ scan_offsets, int n_rings, int cap, FeatArrays A) {
  const int b = blockIdx.x / n_rings, ring = blockIdx.x % n_rings;
  const int out0 = scan_offsets[b] + 8 * b;
  const int ss = A.scan_start[b * kMaxRings + ring], se = A.scan_end[b * kMaxRings + ring];
  const int seg0 = (b * n_rings + ring) * 6;
  if (threadIdx.x < 6 * 5) A.seg_counts[seg0 * 5 + threadIdx.x] = 0;
  if (se - ss < 10) return;  // :471
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem S(smem_raw, cap);
  const float4* C = A.cloud + out0;
  int* NP = A.neighbor_picked + out0;
  int* INP = A.inten_neighbor_picked + out0;
  int* LB = A.label + out0;
  int* ILB = A.inten_label + out0;

  for (int j = 0; j < 6; j++) {
    const int sp = ss + (se - ss) * j / 6;
    const int ep = ss + (se - ss) * (j + 1) / 6 - 1;
    const int len = ep - sp + 1;
    if (len <= 0) continue;  // uniform across the block
    const int w0 = sp - 5, wn = len + 10;
    for (int t = threadIdx.x; t < wn; t += blockDim.x) {
      const int g = w0 + t;
      S.pt[t] = C[g];
      S.cv[t] = A.curvature[out0 + g];
      S.c2[t] = A.curvature2[out0 + g];
      S.ci[t] = A.inten_curvature[out0 + g];
      S.in[t] = A.intensity_num[out0 + g];
      S.np[t] = (signed char)NP[g];
      S.inp[t] = (signed char)INP[g];
      S.lb[t] = (signed char)LB[g];
      S.ilb[t] = (signed char)ILB[g];
      S.gm[t] = (signed char)A.ground_marked[out0 + g];
    }
    int npad = 1;
    while (npad < len) npad <<= 1;
    __syncthreads();
    for (int t = threadIdx.x; t < npad; t += blockDim.x) {
      const bool in = t < len;
      S.k1[t] = in ? S.cv[t + 5] : INFINITY;
      S.i1[t] = in ? t + 5 : 0x7fffffff;  // window-relative index (same order as the global index)
      S.k2[t] = in ? S.ci[t + 5] : INFINITY;
      S.i2[t] = in ? t + 5 : 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort(S.k1, S.i1, npad);
    bitonic_sort(S.k2, S.i2, npad);
    if (threadIdx.x == 0) {
      const int seg = seg0 + j;
      int* cnt = A.seg_counts + seg * 5;
      auto gap2 = [&](int a, int bb) {
        const float4 pa = S.pt[a], pb = S.pt[bb];
        const float dx = fsub(pa.x, pb.x), dy = fsub(pa.y, pb.y), dz = fsub(pa.z, pb.z);
        return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
      };
      auto suppress = [&](int ind) {  // :517-534 / :564-581
        S.np[ind] = 1;
        for (int l = 1; l <= 5; l++) {
          if ((double)gap2(ind + l, ind + l - 1) > 0.05) break;
          S.np[ind + l] = 1;
        }
        for (int l = -1; l >= -5; l--) {
          if ((double)gap2(ind + l, ind + l + 1) > 0.05) break;
          S.np[ind + l] = 1;
        }
      };
      // sharp / less sharp (:485-536)
      int largest = 0, n_sharp = 0, n_less = 0;
      for (int k = len - 1; k >= 0; k--) {
        const int ind = S.i1[k];
        if (S.np[ind] == 0 && S.gm[ind] != 1 && (double)S.cv[ind] > 0.1 && (double)S.c2[ind] > 0.3) {
          largest++;
          if (largest <= 20) {
            S.lb[ind] = 2;
            A.seg_sharp[seg * 20 + n_sharp++] = w0 + ind;
            A.seg_less[seg * 22 + n_less++] = w0 + ind;
          } else if (largest <= 21) {
            S.lb[ind] = 1;
            A.seg_less[seg * 22 + n_less++] = w0 + ind;
          } else {
            break;
          }
          suppress(ind);
        }
      }
      // flat (:538-583)
      int smallest = 0, n_flat = 0;
      for (int k = 0; k < len; k++) {
        const int ind = S.i1[k];
        if (S.np[ind] == 0 && (double)S.cv[ind] < 0.3 && (double)S.c2[ind] < 0.4) {
          smallest++;
          if (smallest <= 40) {
            S.lb[ind] = -1;
            A.seg_flat[seg * 40 + n_flat++] = w0 + ind;
          } else {
            break;
          }
          suppress(ind);
        }
      }
      // intensity edges (:594-641)
      int largest2 = 0, n_inten = 0, n_inten_less = 0;
      for (int k = len - 1; k >= 0; k--) {
        const int ind = S.i2[k];
        if (S.inp[ind] == 0 && S.gm[ind] != 1 && S.ci[ind] > 65 && S.lb[ind] != 2 && S.lb[ind] != 1) {
          largest2++;
          if (largest2 <= 20) {
            S.ilb[ind] = 2;
            A.seg_inten[seg * 20 + n_inten++] = w0 + ind;
            A.seg_inten_less[seg * 21 + n_inten_less++] = w0 + ind;
          } else if (largest2 <= 21) {
            S.ilb[ind] = 1;
            A.seg_inten_less[seg * 21 + n_inten_less++] = w0 + ind;
            A.seg_less[seg * 22 + n_less++] = w0 + ind;  // cornerPointsLessSharp (:617)
          } else {
            break;
          }
          S.inp[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            const float dI = (float)(S.in[ind + l] - S.in[ind + l - 1]);
            if (absf(dI) > 35) break;
            S.inp[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            const float dI = (float)(S.in[ind + l] - S.in[ind + l + 1]);
            if (absf(dI) > 35) break;
            S.inp[ind + l] = 1;
          }
        }
      }
      cnt[0] = n_sharp; cnt[1] = n_less; cnt[2] = n_flat; cnt[3] = n_inten; cnt[4] = n_inten_less;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < wn; t += blockDim.x) {
      const int g = w0 + t;
      NP[g] = S.np[t];
      INP[g] = S.inp[t];
      LB[g] = S.lb[t];
      ILB[g] = S.ilb[t];
    }
    __syncthreads();
  }
}

// One block per scan: per-segment pick buffers -> lists in push_back order (:469-663)
__global__ void __launch_bounds__(128) k_feat_compact(int n_rings, FeatArrays A) {
  const int b = blockIdx.x, nseg = n_rings * 6, seg0 = b * nseg;
  __shared__ int s_off[5][kMaxRings * 6 + 1];
  if (threadIdx.x < 5) {
    int acc = 0;
    for (int s = 0; s < nseg; s++) {
      s_off[threadIdx.x][s] = acc;
      acc += A.seg_counts[(seg0 + s) * 5 + threadIdx.x];
    }
    s_off[threadIdx.x][nseg] = acc;
    A.list_counts[b * 5 + threadIdx.x] = acc;
  }
  __syncthreads();
  const int caps[5] = {20, 22, 40, 20, 21};
  const int* src[5] = {A.seg_sharp, A.seg_less, A.seg_flat, A.seg_inten, A.seg_inten_less};
  int* dst[5] = {A.corner_sharp + (size_t)b * RGC_FEAT_CAP_SHARP(n_rings), A.corner_less_sharp + (size_t)b * RGC_FEAT_CAP_LESS_SHARP(n_rings),
                 A.surf_flat + (size_t)b * RGC_FEAT_CAP_FLAT(n_rings), A.inten_sharp + (size_t)b * RGC_FEAT_CAP_INTEN(n_rings),
                 A.inten_less_sharp + (size_t)b * RGC_FEAT_CAP_LESS_INTEN(n_rings)};
  // (the per-feature weights are filled by k_feat_weights from the scan's per-point arrays)
  for (int s = threadIdx.x; s < nseg; s += blockDim.x)
    for (int L = 0; L < 5; L++) {
      const int c = A.seg_counts[(seg0 + s) * 5 + L];
      for (int e = 0; e < c; e++) dst[L][s_off[L][s] + e] = src[L][(size_t)(seg0 + s) * caps[L] + e];
    }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double sharp = (double)s_off[0][nseg], plane = (double)s_off[2][nseg];
    A.inten_merged[b] = (sharp / plane < 0.3) ? 1 : 0;  // :650-656 (NaN / inf compare false)
  }
}

// weights of the compacted lists (normal_x fields :501,:554,:609); one thread per list entry
__global__ void __launch_bounds__(256) k_feat_weights(const int* __restrict__ scan_offsets, int n_rings, int n_scans, FeatArrays A) {
  const int b = blockIdx.y;
  if (b >= n_scans) return;
  const int out0 = scan_offsets[b] + 8 * b;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = A.list_counts[b * 5 + 0], nf = A.list_counts[b * 5 + 2], ni = A.list_counts[b * 5 + 3];
  if (t < ns) {
    const size_t o = (size_t)b * RGC_FEAT_CAP_SHARP(n_rings) + t;
    A.corner_sharp_w[o] = fadd(A.distance_source[out0 + A.corner_sharp[o]], 1.f);
  }
  if (t < nf) {
    const size_t o = (size_t)b * RGC_FEAT_CAP_FLAT(n_rings) + t;
    A.surf_flat_w[o] = A.distance_source[out0 + A.surf_flat[o]];
  }
  if (t < ni) {
    const size_t o = (size_t)b * RGC_FEAT_CAP_INTEN(n_rings) + t;
    A.inten_sharp_w[o] = A.other_source[out0 + A.inten_sharp[o]];
  }
}

// exclusive prefix sum of one int per thread over a 256-thread block; `total` = the block's sum
__device__ __forceinline__ int block_excl_scan256(int v, int& total) {
  __shared__ int ws[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  __syncthreads();  // ws may still be read by the previous call
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const int sw = ws[w];
    if (w < warp) before += sw;
    tot += sw;
  }
  total = tot;
  return before + x - v;
}

// One block per scan: the two clouds of the reference that are not bounded feature lists.
//  * surfPointsLessFlatScan (:586-592): after the picks of a sextant, every point k of it with cloudLabel <= 0, in
//    ascending k; the sextants of a ring tile [scanStartInd, scanEndInd) and only rings with at least 10 such
//    points are processed (:471), so the list is the ascending k of those ranges with a final label <= 0.
//  * GroundPoints (:338): every (ring, column, n) sample of the ground marking in loop order — duplicates
//    included, this is the cloud published on /laser_cloud_ground (:714-717).  ground_size[b] is its length;
//    entries beyond ground_cap are dropped.
__global__ void __launch_bounds__(256) k_feat_clouds(const int* __restrict__ scan_offsets, int n_rings, FeatArrays A) {
  const int b = blockIdx.x, out0 = scan_offsets[b] + 8 * b;
  const int* sst = A.scan_start + b * kMaxRings;
  const int* sen = A.scan_end + b * kMaxRings;
  if (A.surf_less_flat) {
    const int* LB = A.label + out0;
    int* dst = A.surf_less_flat + out0;
    int base = 0;
    for (int ring = 0; ring < n_rings; ring++) {
      const int ss = sst[ring], se = sen[ring];
      if (se - ss < 10) continue;  // :471 (the six sextants of a processed ring tile [ss, se) exactly)
      for (int k0 = ss; k0 < se; k0 += 256) {
        const int k = k0 + threadIdx.x;
        const int f = (k < se && LB[k] <= 0) ? 1 : 0;
        int total;
        const int off = block_excl_scan256(f, total);
        if (f) dst[base + off] = k;
        base += total;
      }
    }
    if (threadIdx.x == 0) A.n_surf_less_flat[b] = base;
  }
  if (A.ground_points) {
    const float Ground_scan_range[8] = {2.66f, 3.04f, 3.56f, 4.30f, 5.44f, 7.41f, 11.63f, 27.12f};
    const int groundScanInd = 7;
    const float4* C = A.cloud + out0;
    const float* R = A.range_vec + out0;
    int* dst = A.ground_points + (size_t)b * A.ground_cap;
    int base = 0;
    for (int i = 0; i < groundScanInd && i < n_rings; i++) {  // same enumeration as k_feat_ground
      const int ring0 = sst[i] - 5, ring_n = sen[i] + 5 - ring0;
      if (ring_n < 5) continue;
      const float th = (float)(0.8 * (1.0 + i / (groundScanInd - 1)));
      for (int c0 = 5; c0 < ring_n - 5; c0 += 256) {
        const int col = c0 + threadIdx.x;
        unsigned mask = 0;  // bit (n + 5): sample (col, n) is appended
        int ci = 0;
        if (col < ring_n - 5) {
          ci = ring0 + col;
          const float rc = R[ci];
          if (absf(fsub(rc, Ground_scan_range[i])) < th && (double)C[ci].z < 0.3)
            for (int n = -5; n < 5; n++)
              if (absf(fsub(R[ci + n], rc)) < fmul(th, 0.5f)) mask |= 1u << (n + 5);
        }
        int total;
        int off = base + block_excl_scan256(__popc(mask), total);
        for (int n = -5; n < 5; n++)
          if ((mask >> (n + 5)) & 1u) {
            if (off < A.ground_cap) dst[off] = ci + n;
            off++;
          }
        base += total;
      }
    }
  }
}

template <class T>
int alloc_dev(rgc_ctx* c, T*& p, size_t count, std::vector<void*>& owned) {
  p = (T*)c->get(sizeof(T) * (count ? count : 1));
  if (!p) {
    c->err = "device allocation failed (features)";
    return RGC_ERR_NOMEM;
  }
  owned.push_back(p);
  return RGC_OK;
}

}  // namespace
}  // namespace rgc

using namespace rgc;

extern "C" int rgc_feat_extract(rgc_ctx* c, const rgc_scan_batch* batch, rgc_feat_out* out) {
  if (!c || !batch || !out || !batch->xyzi || !batch->scan_offsets || batch->n_scans <= 0) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  const int nb = batch->n_scans, nr = batch->n_rings;
  if (nr != 16 && nr != 32 && nr != 64) FAIL(c, RGC_ERR_UNSUPPORTED, "only 16, 32 or 64 scan lines are supported (scanRegistration.cpp:69)");
  const int total_in = batch->scan_offsets[nb];
  if (total_in <= 0) FAIL(c, RGC_ERR_INVALID, "empty scan batch");
  const size_t total_out = (size_t)total_in + 8 * (size_t)nb;
  const size_t nseg = (size_t)nb * nr * 6;
  cudaStream_t st = c->stream;
  std::vector<void*> owned;
  FeatArrays A{};
  float4* d_raw;
  int *d_off, *tmp_ring, *d_err;
  float* tmp_inten;
#define AL(ptr, count) TRY(alloc_dev(c, ptr, count, owned))
  AL(d_raw, total_in); AL(d_off, nb + 1); AL(tmp_ring, total_in); AL(tmp_inten, total_in); AL(d_err, 1);
  AL(A.cloud, total_out); AL(A.src_index, total_out); AL(A.intensity2, total_out); AL(A.intensity_num, total_out);
  AL(A.range_vec, total_out); AL(A.scan_angle, total_out); AL(A.curvature, total_out); AL(A.inten_curvature, total_out);
  AL(A.curvature2, total_out); AL(A.distance_source, total_out); AL(A.other_source, total_out);
  AL(A.label, total_out); AL(A.inten_label, total_out); AL(A.neighbor_picked, total_out); AL(A.inten_neighbor_picked, total_out);
  AL(A.ground_marked, total_out);
  AL(A.cloud_size, nb); AL(A.scan_start, (size_t)nb * kMaxRings); AL(A.scan_end, (size_t)nb * kMaxRings); AL(A.ground_size, nb);
  AL(A.inten_merged, nb); AL(A.groundparam, (size_t)nb * 11);
  AL(A.seg_sharp, nseg * 20); AL(A.seg_less, nseg * 22); AL(A.seg_flat, nseg * 40); AL(A.seg_inten, nseg * 20); AL(A.seg_inten_less, nseg * 21);
  AL(A.seg_counts, nseg * 5);
  AL(A.corner_sharp, (size_t)nb * RGC_FEAT_CAP_SHARP(nr)); AL(A.corner_sharp_w, (size_t)nb * RGC_FEAT_CAP_SHARP(nr));
  AL(A.corner_less_sharp, (size_t)nb * RGC_FEAT_CAP_LESS_SHARP(nr));
  AL(A.surf_flat, (size_t)nb * RGC_FEAT_CAP_FLAT(nr)); AL(A.surf_flat_w, (size_t)nb * RGC_FEAT_CAP_FLAT(nr));
  AL(A.inten_sharp, (size_t)nb * RGC_FEAT_CAP_INTEN(nr)); AL(A.inten_sharp_w, (size_t)nb * RGC_FEAT_CAP_INTEN(nr));
  AL(A.inten_less_sharp, (size_t)nb * RGC_FEAT_CAP_LESS_INTEN(nr));
  AL(A.list_counts, (size_t)nb * 5);
  A.surf_less_flat = A.n_surf_less_flat = A.ground_points = nullptr;
  A.ground_cap = out->ground_points ? std::max(out->ground_cap, 0) : 0;
  if (out->surf_less_flat || out->n_surf_less_flat) {
    AL(A.surf_less_flat, total_out);
    AL(A.n_surf_less_flat, nb);
  }
  if (A.ground_cap > 0) AL(A.ground_points, (size_t)nb * A.ground_cap);
#undef AL
  auto release = [&]() {
    for (void* p : owned) c->put(p);
  };
  CK(c, cudaMemcpyAsync(d_raw, batch->xyzi, sizeof(float4) * (size_t)total_in, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(d_off, batch->scan_offsets, sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, st));
  CK(c, cudaEventRecord(c->ev[0], st));
  // everything the reference zero-initialises per point (:297-305) plus the arrays it leaves stale
  for (void* p : {(void*)A.scan_angle, (void*)A.curvature, (void*)A.inten_curvature, (void*)A.curvature2, (void*)A.distance_source, (void*)A.other_source,
                  (void*)A.range_vec})
    CK(c, cudaMemsetAsync(p, 0, 4 * total_out, st));
  for (void* p : {(void*)A.label, (void*)A.inten_label, (void*)A.neighbor_picked, (void*)A.inten_neighbor_picked, (void*)A.ground_marked, (void*)A.intensity2,
                  (void*)A.intensity_num, (void*)A.src_index})
    CK(c, cudaMemsetAsync(p, 0, 4 * total_out, st));
  CK(c, cudaMemsetAsync(A.cloud, 0, sizeof(float4) * total_out, st));
  CK(c, cudaMemsetAsync(d_err, 0, 4, st));

  const int pt_blocks = (int)((total_out + 255) / 256);
  k_feat_rings<<<nb, 256, 0, st>>>(d_raw, d_off, nr, (float)batch->minimum_range, (float)batch->maximum_range, tmp_ring, tmp_inten, d_err, A);
  CKL(c);
  int max_seg = 0;  // read back while the per-point passes run
  CK(c, cudaMemcpyAsync(&max_seg, d_err, 4, cudaMemcpyDeviceToHost, st));
  k_feat_pass_a<<<pt_blocks, 256, 0, st>>>(d_off, nb, (int)total_out, A);
  CKL(c);
  k_feat_pass_b<<<pt_blocks, 256, 0, st>>>(d_off, nb, (int)total_out, A);
  CKL(c);
  k_feat_ground<<<nb, 256, 0, st>>>(d_off, nr, A);
  CKL(c);
  k_feat_occlusion<<<pt_blocks, 256, 0, st>>>(d_off, nb, (int)total_out, A);
  CKL(c);
  CK(c, cudaStreamSynchronize(st));  // max_seg (one 4-byte readback per batch)
  if (max_seg > kSegCap) {
    release();
    FAIL(c, RGC_ERR_UNSUPPORTED, "a ring sextant is longer than 2048 points");
  }
  int cap = 64;
  while (cap < max_seg) cap <<= 1;
  const size_t sel_smem = sel_smem_bytes(cap);
  CK(c, cudaFuncSetAttribute(k_feat_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  k_feat_select<<<nb * nr, kSelThreads, sel_smem, st>>>(d_off, nr, cap, A);
  CKL(c);
  k_feat_compact<<<nb, 128, 0, st>>>(nr, A);
  CKL(c);
  {
    dim3 grid((RGC_FEAT_CAP_FLAT(nr) + 255) / 256, nb);
    k_feat_weights<<<grid, 256, 0, st>>>(d_off, nr, nb, A);
    CKL(c);
  }
  if (A.surf_less_flat || A.ground_points) {
    k_feat_clouds<<<nb, 256, 0, st>>>(d_off, nr, A);
    CKL(c);
  }
  CK(c, cudaEventRecord(c->ev[1], st));

  auto d2h = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
  };
  cudaError_t e = cudaSuccess;
#define CP(field, src, count) if (e == cudaSuccess) e = d2h(out->field, src, sizeof(*out->field) * (count))
  CP(cloud_size, A.cloud_size, nb); CP(scan_start, A.scan_start, (size_t)nb * kMaxRings); CP(scan_end, A.scan_end, (size_t)nb * kMaxRings);
  CP(groundparam, A.groundparam, (size_t)nb * 11); CP(ground_size, A.ground_size, nb); CP(inten_merged, A.inten_merged, nb);
  if (e == cudaSuccess) e = d2h(out->cloud, A.cloud, sizeof(float4) * total_out);
  CP(src_index, A.src_index, total_out); CP(intensity_num, A.intensity_num, total_out); CP(range_vec, A.range_vec, total_out);
  CP(scan_angle, A.scan_angle, total_out); CP(curvature, A.curvature, total_out); CP(inten_curvature, A.inten_curvature, total_out);
  CP(curvature2, A.curvature2, total_out); CP(distance_source, A.distance_source, total_out); CP(other_source, A.other_source, total_out);
  CP(label, A.label, total_out); CP(inten_label, A.inten_label, total_out); CP(neighbor_picked, A.neighbor_picked, total_out);
  CP(inten_neighbor_picked, A.inten_neighbor_picked, total_out); CP(ground_marked, A.ground_marked, total_out);
  CP(corner_sharp, A.corner_sharp, (size_t)nb * RGC_FEAT_CAP_SHARP(nr)); CP(corner_sharp_w, A.corner_sharp_w, (size_t)nb * RGC_FEAT_CAP_SHARP(nr));
  CP(corner_less_sharp, A.corner_less_sharp, (size_t)nb * RGC_FEAT_CAP_LESS_SHARP(nr));
  CP(surf_flat, A.surf_flat, (size_t)nb * RGC_FEAT_CAP_FLAT(nr)); CP(surf_flat_w, A.surf_flat_w, (size_t)nb * RGC_FEAT_CAP_FLAT(nr));
  CP(inten_sharp, A.inten_sharp, (size_t)nb * RGC_FEAT_CAP_INTEN(nr)); CP(inten_sharp_w, A.inten_sharp_w, (size_t)nb * RGC_FEAT_CAP_INTEN(nr));
  CP(inten_less_sharp, A.inten_less_sharp, (size_t)nb * RGC_FEAT_CAP_LESS_INTEN(nr));
  if (A.surf_less_flat) {
    CP(surf_less_flat, A.surf_less_flat, total_out);
    CP(n_surf_less_flat, A.n_surf_less_flat, nb);
  }
  if (A.ground_points) CP(ground_points, A.ground_points, (size_t)nb * A.ground_cap);
#undef CP
  // list counts: [scan][5] -> five arrays
  std::vector<int> counts((size_t)nb * 5);
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts.data(), A.list_counts, sizeof(int) * counts.size(), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  release();
  if (e != cudaSuccess) {
    c->err = std::string("rgc_feat_extract: ") + cudaGetErrorString(e);
    return RGC_ERR_CUDA;
  }
  int* lists[5] = {out->n_corner_sharp, out->n_corner_less_sharp, out->n_surf_flat, out->n_inten_sharp, out->n_inten_less_sharp};
  for (int L = 0; L < 5; L++)
    if (lists[L])
      for (int b = 0; b < nb; b++) lists[L][b] = counts[(size_t)b * 5 + L];
  CK(c, cudaEventElapsedTime(&out->device_ms, c->ev[0], c->ev[1]));
  return RGC_OK;
}
