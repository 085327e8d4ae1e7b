// rgc_batch.cuh — kernels of the batched registration path (include/rgc_batch.h).
//
// B independent (source, target) pairs live in two multi-cloud grids (rgc_grid.cuh: CloudRange): pair p's
// source is the sorted positions [src_lo, src_hi) of the combined source cloud, its target the positions
// [tgt_lo, tgt_hi) of the combined target cloud.  One launch serves every pair that is still iterating:
//   k_bcorrespond     update_correspondences, search part      (fast_gicp_impl.hpp:115-137)
//   k_blinearize      Mahalanobis + linearize                   (:139-211)
//   k_bcompute_error  compute_error                             (:214-237)
//   k_bfitness        pcl::Registration::getFitnessScore
// A block belongs to ONE pair and plays the role of one block of the single-registration launch for that
// pair (same 128 points per block, same grid-stride, same warp-shuffle tree, same final pass over the
// pair's partials: grid_reduce_at), so err / H / b of a pair are bit-identical to rgc_reg_linearize on it.
#pragma once
#include "rgc_kernels.cuh"

namespace rgc {

struct BPairInfo {  // static for the life of a chunk
  int src_lo, src_hi;
  int tgt_lo, tgt_hi;
  unsigned long long tgt_prefix;
  int blk0, nblk;    // the pair's blocks in the linearize-shaped launches: nblk = reduce_grid(n_source)
  int fblk0, fnblk;  // ... in the fitness launch: div_up(n_source * spread, kThreads)
  int spread, pad;
};

struct BPairRound {  // uploaded before every LM round
  double T[12];      // pose of this round's linearize (and compute_error), row-major 3x4
  float Tf[12];      // the same cast to float (Eigen::Isometry3d::cast<float>, fast_gicp_impl.hpp:119)
  int active;        // 0: the pair has finished, its blocks exit at once
  int do_ce;         // compute_error at T on buffer set rsel (an LM trial); 0 for the first linearize and for GN
  int rsel, wsel;    // buffer sets read by compute_error / written by linearize
  int hint_sel;      // buffer set whose correspondences seed the search, -1 = none
  int want_hb;
};

struct BCloudSrc {  // where cloud c's raw records sit in the staging buffer
  unsigned long long byte_off;
  unsigned int stride, pad;
};

// raw records of all clouds of a chunk (any PCL stride each) -> one float4 array, cloud after cloud
__global__ void __launch_bounds__(256) k_bingest(const unsigned char* __restrict__ raw, const BCloudSrc* __restrict__ cs, const int* __restrict__ cloud_off, int n_clouds,
                                                 int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = 0, b = n_clouds;
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    if (cloud_off[mid] <= i) a = mid; else b = mid;
  }
  const BCloudSrc s = cs[a];
  const float* p = reinterpret_cast<const float*>(raw + s.byte_off + (size_t)(i - cloud_off[a]) * s.stride);
  out[i] = make_float4(p[0], p[1], p[2], 1.0f);
}

__global__ void __launch_bounds__(kThreads, RGC_CORR_MINB) k_bcorrespond(GridView tgt, const float4* __restrict__ src, const BPairInfo* __restrict__ info,
                                                                        const int* __restrict__ blk_pair, const BPairRound* __restrict__ rounds, float thr2,
                                                                        int* corr0, int* corr1, float* __restrict__ sqd0, float* __restrict__ sqd1,
                                                                        int* __restrict__ need_state, int* __restrict__ need_list, int* __restrict__ need_count) {
  const int pair = blk_pair[blockIdx.x];
  const BPairRound* pr = rounds + pair;
  if (!pr->active) return;
  const BPairInfo pi = info[pair];
  const int vb = blockIdx.x - pi.blk0;
  float Tf[12];
#pragma unroll
  for (int j = 0; j < 12; j++) Tf[j] = pr->Tf[j];
  const int wsel = pr->wsel, hsel = pr->hint_sel;
  int* corr = wsel ? corr1 : corr0;
  float* sqd = wsel ? sqd1 : sqd0;
  const int* hint = hsel < 0 ? nullptr : (hsel ? corr1 : corr0);
  const CloudRange cr{pi.tgt_lo, pi.tgt_hi, pi.tgt_prefix};
  const int n_src = pi.src_hi - pi.src_lo;
  for (int il = vb * kThreads + threadIdx.x; il < n_src; il += pi.nblk * kThreads) {
    const int i = pi.src_lo + il;
    const float4 p = __ldg(&src[i]);
    float qx, qy, qz;
    transform_f(Tf, p.x, p.y, p.z, qx, qy, qz);
    Best1 top;
    top.reset(1, thr2);
    knn_search(tgt, qx, qy, qz, 1, thr2, hint ? hint[i] : -1, top, nullptr, &cr);
    const int pos = (top.id0 >= 0 && top.d0 < thr2) ? top.id0 : -1;
    corr[i] = pos;
    sqd[i] = top.d0;
    if (need_state) {
      const bool claim = pos >= 0 && need_state[pos] == 0 && atomicCAS(&need_state[pos], 0, 1) == 0;
      const unsigned act = __activemask();
      const unsigned m = __ballot_sync(act, claim);
      if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(need_count, __popc(m));
        base = __shfl_sync(act, base, leader);
        if (claim) need_list[base + __popc(m & ((1u << lane) - 1u))] = pos;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, RGC_LIN_MINB) k_blinearize(const float4* __restrict__ tgt_pts, const float4* __restrict__ src, const double* __restrict__ src_cov,
                                                            const double* __restrict__ tgt_cov, const BPairInfo* __restrict__ info, const int* __restrict__ blk_pair,
                                                            const BPairRound* __restrict__ rounds, const int* __restrict__ corr0, const int* __restrict__ corr1,
                                                            double* __restrict__ maha0, double* __restrict__ maha1, double* __restrict__ partials,
                                                            unsigned int* __restrict__ tickets, double* __restrict__ results) {
  const int pair = blk_pair[blockIdx.x];
  const BPairRound* pr = rounds + pair;
  if (!pr->active) return;
  const BPairInfo pi = info[pair];
  const int vb = blockIdx.x - pi.blk0;
  Rt Td;
#pragma unroll
  for (int j = 0; j < 12; j++) Td.m[j] = pr->T[j];
  const int want_hb = pr->want_hb;
  const int* corr = pr->wsel ? corr1 : corr0;
  double* maha = pr->wsel ? maha1 : maha0;
  double acc[kLinN];
#pragma unroll
  for (int j = 0; j < kLinN; j++) acc[j] = 0.0;
  linearize_points(tgt_pts, src, src_cov, tgt_cov, corr, maha, Td, want_hb, pi.src_lo, vb * kThreads + threadIdx.x, pi.nblk * kThreads, pi.src_hi - pi.src_lo, acc);
  grid_reduce_at<kLinN>(acc, partials + (size_t)pi.blk0 * kLinN, (unsigned)vb, (unsigned)pi.nblk, tickets + pair, results + (size_t)pair * 32);
}

__global__ void __launch_bounds__(kThreads) k_bcompute_error(const float4* __restrict__ tgt_pts, const float4* __restrict__ src, const BPairInfo* __restrict__ info,
                                                             const int* __restrict__ blk_pair, const BPairRound* __restrict__ rounds, const int* __restrict__ corr0,
                                                             const int* __restrict__ corr1, const double* __restrict__ maha0, const double* __restrict__ maha1,
                                                             double* __restrict__ partials, unsigned int* __restrict__ tickets, double* __restrict__ results) {
  const int pair = blk_pair[blockIdx.x];
  const BPairRound* pr = rounds + pair;
  if (!pr->active || !pr->do_ce) return;
  const BPairInfo pi = info[pair];
  const int vb = blockIdx.x - pi.blk0;
  Rt Td;
#pragma unroll
  for (int j = 0; j < 12; j++) Td.m[j] = pr->T[j];
  const int* corr = pr->rsel ? corr1 : corr0;
  const double* maha = pr->rsel ? maha1 : maha0;
  double acc[1] = {0.0};
  compute_error_points(tgt_pts, src, corr, maha, Td, pi.src_lo, vb * kThreads + threadIdx.x, pi.nblk * kThreads, pi.src_hi - pi.src_lo, acc);
  grid_reduce_at<1>(acc, partials + (size_t)pi.blk0, (unsigned)vb, (unsigned)pi.nblk, tickets + pair, results + pair);
}

// getFitnessScore of every pair at its final transformation: the thread layout of k_fitness (sparse warps
// for small clouds included) replayed per pair, so the sums come out bit-identical
// final_Tf: 16 floats per pair — the final transformation (12) and, as int bits, the buffer set holding the
// pair's last correspondences (hints, see k_fitness) or -1
// getFitnessScore, search half: the exact nearest target point of every source point at the pair's final pose, one lane
// per query over the linearize-shaped grid (the batch has no idle lanes to give a query), hinted by the pair's last
// correspondences.  d2 (or -1: no target point at all) goes to the sqd buffer of the set the pair is NOT using.
__global__ void __launch_bounds__(kThreads, RGC_CORR_MINB) k_bfitness_search(GridView tgt, const float4* __restrict__ src, const BPairInfo* __restrict__ info,
                                                                            const int* __restrict__ blk_pair, const float* __restrict__ final_Tf,
                                                                            const int* __restrict__ corr0, const int* __restrict__ corr1, float* __restrict__ sqd0,
                                                                            float* __restrict__ sqd1) {
  const int pair = blk_pair[blockIdx.x];
  const BPairInfo pi = info[pair];
  const int vb = blockIdx.x - pi.blk0;
  float Tf[12];
#pragma unroll
  for (int j = 0; j < 12; j++) Tf[j] = final_Tf[pair * 16 + j];
  const int hsel = __float_as_int(final_Tf[pair * 16 + 12]);
  const int* hint = hsel < 0 ? nullptr : (hsel ? corr1 : corr0);
  float* out = hsel == 1 ? sqd0 : sqd1;
  const CloudRange cr{pi.tgt_lo, pi.tgt_hi, pi.tgt_prefix};
  const int n_src = pi.src_hi - pi.src_lo;
  for (int il = vb * kThreads + threadIdx.x; il < n_src; il += pi.nblk * kThreads) {
    const int i = pi.src_lo + il;
    const float4 p = __ldg(&src[i]);
    float qx, qy, qz;
    transform_f(Tf, p.x, p.y, p.z, qx, qy, qz);
    Best1 top;
    top.reset(1, INFINITY);
    knn_search(tgt, qx, qy, qz, 1, INFINITY, hint ? __ldg(&hint[i]) : -1, top, nullptr, &cr);
    out[i] = top.id0 >= 0 ? top.d0 : -1.0f;
  }
}

// getFitnessScore, reduction half: [sum d2, count] over d2 <= max_range in the thread layout and summation order of the
// single registration's k_fitness (a query's value sits on the first of its `spread` lanes), hence the same bits
__global__ void __launch_bounds__(kThreads) k_bfitness(const BPairInfo* __restrict__ info, const int* __restrict__ fblk_pair, const float* __restrict__ final_Tf,
                                                      double max_range, const float* __restrict__ sqd0, const float* __restrict__ sqd1,
                                                      double* __restrict__ partials, unsigned int* __restrict__ tickets, double* __restrict__ results) {
  const int pair = fblk_pair[blockIdx.x];
  const BPairInfo pi = info[pair];
  const int vb = blockIdx.x - pi.fblk0;
  const int hsel = __float_as_int(final_Tf[pair * 16 + 12]);
  const float* d2 = hsel == 1 ? sqd0 : sqd1;
  double acc[2] = {0.0, 0.0};
  const int gt = vb * kThreads + threadIdx.x;
  const int il = gt / pi.spread;
  if ((gt & (pi.spread - 1)) == 0 && il < pi.src_hi - pi.src_lo) {
    const float d0 = d2[pi.src_lo + il];
    if (d0 >= 0.f && (double)d0 <= max_range) {
      acc[0] = (double)d0;
      acc[1] = 1.0;
    }
  }
  grid_reduce_at<2>(acc, partials + (size_t)pi.fblk0 * 2, (unsigned)vb, (unsigned)pi.fnblk, tickets + pair, results + (size_t)pair * 2);
}

}  // namespace rgc
