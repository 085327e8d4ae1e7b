// rgc_kernels.cuh — sm_100a kernels of the scan-matching path (included by rgc_gicp.cu only).
//
// Kernel inventory (each cites the reference code it replaces):
//   k_ingest          AoS point cloud (any PCL stride) -> float4 SoA + per-block bbox partials
//   k_morton          Morton key of the finest voxel of every point
//   k_rs_hist / k_rs_scan / k_rs_scatter   LSD radix sort (8-bit digits, stable)
//   k_gather_sorted   points in Morton order, original index packed in .w
//   k_count_cells / k_build_tables         per-level voxel hash tables over the sorted array
//   k_knn             exact kNN (pcl::search::KdTree::nearestKSearch; fast_gicp_impl.hpp:133,254)
//   k_covariance      k-neighbour covariance + regularisation (fast_gicp_impl.hpp:256-293)
//   k_linearize       1-NN + Mahalanobis + H/b/err reduction (fast_gicp_impl.hpp:115-211)
//   k_compute_error   err reduction with frozen correspondences (fast_gicp_impl.hpp:214-237)
//   k_fitness         pcl::Registration::getFitnessScore
//   k_transform_out   pcl::transformPointCloud (lsq_registration_impl.hpp:78)
#pragma once
#include <cuda_runtime.h>

#include "rgc_grid.cuh"
#include "rgc_math.cuh"

namespace rgc {

// debug statistics buffer (rgc_debug_*_stats); null in normal operation
__device__ long long* g_tile_dbg = nullptr;

constexpr int kBboxBlocks = 296;  // 2 x 148 SMs
constexpr int kThreads = 128;

// Small clouds (one LiDAR sweep) cannot fill 148 SMs with one query per thread: the search kernels
// are then bound by the latency of a few divergent warps.  `spread` (power of two) gives every query
// `spread` consecutive lanes.  Round 1 left the extra lanes idle (fewer divergent walks per warp; measured on
// 22k queries with one walking lane: spread 1: 137 us, 4: 92 us, 32: 146 us); since round 2 the lanes of a
// group share the query (nn1_search_group: each lane takes some of the ball's root cells), and 8 lanes per query
// beat 4 on the 22k-point sweep (cold align 0.865 -> 0.812 ms, LM loop 0.62 -> 0.575 ms).
// Large clouds use spread = 1 (throughput-bound, one lane per query).
#ifndef RGC_SPREAD8_MAX
#define RGC_SPREAD8_MAX 48000
#endif
__host__ __device__ inline int query_spread(int n) { return n <= RGC_SPREAD8_MAX ? 8 : (n <= 96000 ? 4 : 1); }


// ------------------------------------------------------------------------------------------------
// ingest: raw[n] with byte stride (xyz at offset 0, as every PCL point type) -> float4(x,y,z,1)
// plus per-block min/max partials (reduced on the host: 296 x 6 floats).
// block reduction of per-thread bbox partials; a block that saw a non-finite coordinate writes NaN
// partials (fminf / fmaxf drop NaN operands, so the flag travels separately) and the host rejects the
// cloud: a NaN candidate would silently break the exactness of the k-NN heaps
__device__ __forceinline__ void bbox_block_reduce(float mn[3], float mx[3], bool bad, float* __restrict__ bbox_partials) {
  __shared__ float sm[8][6];
#pragma unroll
  for (int a = 0; a < 3; a++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int a = 0; a < 3; a++) {
      sm[warp][a] = mn[a];
      sm[warp][3 + a] = mx[a];
    }
  const bool any_bad = __syncthreads_or(bad) != 0;
  if (threadIdx.x < 6) {
    float v = sm[0][threadIdx.x];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = threadIdx.x < 3 ? fminf(v, sm[w][threadIdx.x]) : fmaxf(v, sm[w][threadIdx.x]);
    bbox_partials[blockIdx.x * 6 + threadIdx.x] = any_bad ? NAN : v;
  }
}

__global__ void __launch_bounds__(256) k_ingest(const unsigned char* __restrict__ raw, size_t stride, int n, float4* __restrict__ out,
                                                float* __restrict__ bbox_partials) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
    float x = p[0], y = p[1], z = p[2];
    if (out) out[i] = make_float4(x, y, z, 1.0f);
    bad |= !(isfinite(x) && isfinite(y) && isfinite(z));
    mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
    mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
    mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
  }
  bbox_block_reduce(mn, mx, bad, bbox_partials);
}

struct GridGeom {
  float inv_s0;
  int bias, nbits;
};

// `cloud_off` (nullable, n_clouds + 1 offsets into pts): a multi-cloud grid — point i of cloud c gets
// the key prefix c << 3 * nbits, so the sort groups the clouds and every cell belongs to one cloud
__global__ void __launch_bounds__(256) k_morton(const float4* __restrict__ pts, int n, GridGeom g, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                const int* __restrict__ cloud_off, int n_clouds) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  const int hi = (1 << g.nbits) - 1;
  int cx = min(max(cell_coord(p.x, g.inv_s0, g.bias), 0), hi);
  int cy = min(max(cell_coord(p.y, g.inv_s0, g.bias), 0), hi);
  int cz = min(max(cell_coord(p.z, g.inv_s0, g.bias), 0), hi);
  uint64_t key = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
  if (cloud_off) {
    int a = 0, b = n_clouds;  // cloud_off[a] <= i < cloud_off[b]
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (cloud_off[mid] <= i) a = mid; else b = mid;
    }
    key |= (uint64_t)a << (3 * g.nbits);
  }
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, 8-bit digits.  Tile = 256 threads x RS_ITEMS keys; inside a tile the key order
// is (warp, round, lane) == ascending index, so ranks computed per (warp, round) with match_any
// are stable.  hist is digit-major [256][nblk] so one exclusive scan yields every block's base.
#ifndef RGC_RS_ITEMS
#define RGC_RS_ITEMS 8
#endif
constexpr int RS_ITEMS = RGC_RS_ITEMS;
constexpr int RS_TILE = 256 * RS_ITEMS;

__global__ void __launch_bounds__(256) k_rs_hist(const uint64_t* __restrict__ keys, int n, int shift, uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int i = base + r * 256 + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// one block per digit: exclusive scan of that digit's per-tile counts in place + the digit total.
// (The first version scanned all 256*nblk counters in ONE block: 27 us per pass on a 500k cloud.)
__global__ void __launch_bounds__(256) k_rs_scan(uint32_t* __restrict__ hist, int nblk, uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t warp_sums[8];
  uint32_t* row = hist + (size_t)blockIdx.x * nblk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < nblk; base += 256) {
    const int i = base + threadIdx.x;
    const uint32_t c = i < nblk ? row[i] : 0u;
    uint32_t v = c;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int w = 0; w < 8; w++) {
      const uint32_t s = warp_sums[w];
      if (w < warp) before += s;
      total += s;
    }
    if (i < nblk) row[i] = carry + before + v - c;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                    uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                    const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ digit_total, int n, int shift,
                                                    int nblk) {
  __shared__ uint32_t cnt[8][256];
  __shared__ uint32_t digit_base[256];
  {  // exclusive prefix of the 256 digit totals (every block recomputes it: 256 values)
    __shared__ uint32_t ws[8];
    const uint32_t c = digit_total[threadIdx.x];
    uint32_t v = c;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) >= o) v += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += ws[w];
    digit_base[threadIdx.x] = before + v - c;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_base = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&cnt[0][0])[i] = 0;
  __syncthreads();
  uint64_t k[RS_ITEMS];
  uint32_t d[RS_ITEMS];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int i = warp_base + r * 32 + lane;
    bool valid = i < n;
    k[r] = valid ? keys_in[i] : 0ull;
    d[r] = valid ? ((uint32_t)(k[r] >> shift) & 255u) : (256u + lane);  // invalid lanes never match
    uint32_t peers = __match_any_sync(0xffffffffu, d[r]);
    if (valid && (peers & lt_mask) == 0) cnt[warp][d[r]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {  // per digit: running base over the 8 warps
    uint32_t base = digit_base[threadIdx.x] + offsets[threadIdx.x * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < 8; w++) {
      uint32_t c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = base;
      base += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int i = warp_base + r * 32 + lane;
    bool valid = i < n;
    uint32_t peers = __match_any_sync(0xffffffffu, d[r]);
    uint32_t rank = __popc(peers & lt_mask);
    uint32_t off = valid ? cnt[warp][d[r]] : 0u;
    __syncwarp();
    if (valid && rank == 0) cnt[warp][d[r]] = off + __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[off + rank] = k[r];
      vals_out[off + rank] = vals_in[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Single-pass-per-digit LSD radix sort ("onesweep").  The three launches per digit above (per-tile histogram,
// scan, scatter: ~18 us per digit on the 500k-point submap, most of it launch gaps and half-empty kernels) become
// ONE: the global digit histograms of ALL passes are counted up front (k_keys_hist, fused with the key
// generation, or k_rs_ghist for keys that already exist), and each pass finds its tile's base inside a digit
// by a decoupled look-back over the tiles before it: a tile first publishes its own per-digit counts
// (AGGREGATE), then walks back adding predecessors' counts until it meets one that already knows its
// inclusive PREFIX.  Tile ids come from an atomic counter, so a tile's predecessors have all started and
// publish their aggregates without waiting for anybody: the look-back always terminates.  Same per-warp
// match_any ranking as k_rs_scatter: stable.
// Scratch layout (uint32, zeroed before the first pass): [0, 8) tile counters per pass, [8, 16) error flag + pad,
// [16, 16 + 8 * 256) global digit histograms per pass, then per pass nblk * 256 status words
// (flag << 30 | count; flag 1 = aggregate, 2 = inclusive prefix).
constexpr int RS_MAX_PASSES = 8;
#ifndef RGC_RS_LOOKBACK
#define RGC_RS_LOOKBACK 8
#endif
constexpr int RS_LOOKBACK = RGC_RS_LOOKBACK;  // predecessor tiles read per look-back round trip
constexpr int RS_GHIST_OFF = 16;
constexpr int RS_STATUS_OFF = RS_GHIST_OFF + RS_MAX_PASSES * 256;
__host__ __device__ inline size_t rs_scratch_words(int n, int passes) { return (size_t)RS_STATUS_OFF + (size_t)passes * (size_t)((n + RS_TILE - 1) / RS_TILE) * 256; }

// digit histograms of all passes for existing keys
__global__ void __launch_bounds__(256) k_rs_ghist(const uint64_t* __restrict__ keys, int n, int passes, uint32_t* __restrict__ scratch) {
  __shared__ uint32_t h[RS_MAX_PASSES * 256];
  for (int i = threadIdx.x; i < passes * 256; i += 256) h[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    for (int p = 0; p < passes; p++) atomicAdd(&h[p * 256 + ((uint32_t)(key >> (8 * p)) & 255u)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += 256)
    if (h[i]) atomicAdd(&scratch[RS_GHIST_OFF + i], h[i]);
}

// Morton keys + the digit histograms of all sort passes, optionally fused with the ingest (raw AoS -> float4
// + bounding-box partials) when the grid geometry is known before the bounding box is (speculative build:
// the geometry of the previous cloud of the same stream, verified on the host once the box has arrived)
template <bool INGEST>
__global__ void __launch_bounds__(256) k_keys_hist(const unsigned char* __restrict__ raw, size_t stride, int n, float4* pts, float* __restrict__ bbox_partials,
                                                   GridGeom g, const int* __restrict__ cloud_off, int n_clouds, int passes, uint64_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals, uint32_t* __restrict__ scratch) {
  pdl_enter();
  __shared__ uint32_t h[RS_MAX_PASSES * 256];
  for (int i = threadIdx.x; i < passes * 256; i += 256) h[i] = 0;
  __syncthreads();
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
  const int hi = (1 << g.nbits) - 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float x, y, z;
    if (INGEST) {
      const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
      x = p[0]; y = p[1]; z = p[2];
      pts[i] = make_float4(x, y, z, 1.0f);
      bad |= !(isfinite(x) && isfinite(y) && isfinite(z));
      mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x);
      mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
      mn[2] = fminf(mn[2], z); mx[2] = fmaxf(mx[2], z);
    } else {
      const float4 p = pts[i];
      x = p.x; y = p.y; z = p.z;
    }
    const int cx = min(max(cell_coord(x, g.inv_s0, g.bias), 0), hi);
    const int cy = min(max(cell_coord(y, g.inv_s0, g.bias), 0), hi);
    const int cz = min(max(cell_coord(z, g.inv_s0, g.bias), 0), hi);
    uint64_t key = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
    if (cloud_off) {
      int a = 0, b = n_clouds;  // cloud_off[a] <= i < cloud_off[b]
      while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (cloud_off[mid] <= i) a = mid; else b = mid;
      }
      key |= (uint64_t)a << (3 * g.nbits);
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
    for (int p = 0; p < passes; p++) atomicAdd(&h[p * 256 + ((uint32_t)(key >> (8 * p)) & 255u)], 1u);
  }
  if (INGEST) bbox_block_reduce(mn, mx, bad, bbox_partials);
  __syncthreads();
  for (int i = threadIdx.x; i < passes * 256; i += 256)
    if (h[i]) atomicAdd(&scratch[RS_GHIST_OFF + i], h[i]);
}

// one sort pass.  LAST_GATHER: the final pass also gathers the points into sorted order (float4 with the
// original index in .w) and writes the inverse permutation, instead of the value array.
template <bool LAST_GATHER>
__global__ void __launch_bounds__(256) k_rs_onesweep(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
                                                     uint32_t* __restrict__ vals_out, uint32_t* scratch, int pass, int n, int nblk,
                                                     const float4* __restrict__ pts, float4* __restrict__ sorted, int* __restrict__ inv) {
  pdl_enter();
  __shared__ uint32_t cnt[8][256];
  __shared__ uint32_t digit_base[256];
  __shared__ uint32_t s_tile;
  const int shift = 8 * pass;
  if (threadIdx.x == 0) s_tile = atomicAdd(&scratch[pass], 1u);
  {  // exclusive prefix of this pass's 256 digit totals
    __shared__ uint32_t ws[8];
    const uint32_t c = scratch[RS_GHIST_OFF + pass * 256 + threadIdx.x];
    uint32_t v = c;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) >= o) v += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += ws[w];
    digit_base[threadIdx.x] = before + v - c;
  }
  for (int i = threadIdx.x; i < 8 * 256; i += 256) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_base = (int)tile * RS_TILE + warp * (32 * RS_ITEMS);
  uint64_t k[RS_ITEMS];
  uint32_t d[RS_ITEMS], vv[RS_ITEMS];
  const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int i = warp_base + r * 32 + lane;
    const bool valid = i < n;
    k[r] = valid ? keys_in[i] : 0ull;
    vv[r] = valid ? vals_in[i] : 0u;  // loaded with the keys: one memory round trip less behind the look-back
    d[r] = valid ? ((uint32_t)(k[r] >> shift) & 255u) : (256u + lane);  // invalid lanes never match
    const uint32_t peers = __match_any_sync(0xffffffffu, d[r]);
    if (valid && (peers & lt_mask) == 0) cnt[warp][d[r]] += __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {  // thread = digit: publish the tile's count, look back for the count of all tiles before it
    const uint32_t dg = threadIdx.x;
    uint32_t agg = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) agg += cnt[w][dg];
    volatile uint32_t* status = scratch + RS_STATUS_OFF + (size_t)pass * nblk * 256;
    uint32_t excl = 0;
    if (tile == 0) {
      status[dg] = (2u << 30) | agg;
    } else {
      status[(size_t)tile * 256 + dg] = (1u << 30) | agg;
      // look back RS_LOOKBACK tiles per round trip (independent loads), consuming them nearest first up to the first
      // inclusive prefix; a tile that has not published yet is polled again.  (One tile per round trip made a pass
      // over the 245 tiles of the 500k-point submap 17 us: all tiles are resident at once, so the prefix
      // frontier had to crawl forward one L2 round trip at a time.)
      int p = (int)tile - 1;
      unsigned spins = 0;
      for (;;) {
        uint32_t v[RS_LOOKBACK];
#pragma unroll
        for (int j = 0; j < RS_LOOKBACK; j++) v[j] = p - j >= 0 ? status[(size_t)(p - j) * 256 + dg] : (2u << 30);  // before tile 0: prefix 0
        int used = 0;
        bool stop = false, stall = false;
#pragma unroll
        for (int j = 0; j < RS_LOOKBACK; j++) {
          const uint32_t f = v[j] >> 30;
          if (!stop && !stall) {
            if (f == 0) {
              stall = true;
            } else {
              excl += v[j] & 0x3fffffffu;
              used++;
              stop = f == 2u;
            }
          }
        }
        if (stop) break;
        p -= used;
        if (used == 0) {
          if (++spins > (1u << 22)) {  // a predecessor that never publishes: report instead of hanging (cannot happen, see above)
            scratch[8] = 1u;
            break;
          }
        } else {
          spins = 0;
        }
      }
      status[(size_t)tile * 256 + dg] = (2u << 30) | (excl + agg);
    }
    uint32_t base = digit_base[dg] + excl;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint32_t c = cnt[w][dg];
      cnt[w][dg] = base;
      base += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int i = warp_base + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t peers = __match_any_sync(0xffffffffu, d[r]);
    const uint32_t rank = __popc(peers & lt_mask);
    const uint32_t off = valid ? cnt[warp][d[r]] : 0u;
    __syncwarp();
    if (valid && rank == 0) cnt[warp][d[r]] = off + __popc(peers);
    __syncwarp();
    if (valid) {
      const uint32_t dst = off + rank;
      keys_out[dst] = k[r];
      const uint32_t o = vv[r];
      if (LAST_GATHER) {
        float4 pt = pts[o];
        pt.w = __int_as_float((int)o);
        sorted[dst] = pt;
        inv[o] = (int)dst;
      } else {
        vals_out[dst] = o;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_gather_sorted(const float4* __restrict__ pts, const uint32_t* __restrict__ vals, int n, float4* __restrict__ sorted,
                                                       int* __restrict__ inv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t o = vals[i];
  float4 p = pts[o];
  p.w = __int_as_float((int)o);
  sorted[i] = p;
  inv[o] = i;
}

// cells per level: point i opens a new cell at every level l with 3l <= highest differing bit.
// counts[0 .. kMaxLevels) are the (zeroed) counters, counts[kMaxLevels] a (zeroed) ticket: the last block to
// finish publishes the counters, followed by the radix sort's error word, in mapped pinned host memory.
__global__ void __launch_bounds__(256) k_count_cells(const uint64_t* __restrict__ keys, int n, int nlevels, uint32_t* __restrict__ counts,
                                                     const uint32_t* __restrict__ sort_err, uint32_t* __restrict__ host_out) {
  pdl_enter();
  __shared__ uint32_t c[kMaxLevels];
  __shared__ bool is_last;
  if (threadIdx.x < kMaxLevels) c[threadIdx.x] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int top;  // highest level at which i starts a new cell
    if (i == 0)
      top = nlevels - 1;
    else {
      uint64_t x = keys[i] ^ keys[i - 1];
      top = x ? min((63 - __clzll((long long)x)) / 3, nlevels - 1) : -1;
    }
    for (int l = 0; l <= top; l++) atomicAdd(&c[l], 1u);
  }
  __syncthreads();
  if (threadIdx.x < nlevels && c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], c[threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&counts[kMaxLevels], 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < kMaxLevels) host_out[threadIdx.x] = __ldcg(&counts[threadIdx.x]);
    if (threadIdx.x == kMaxLevels) host_out[kMaxLevels] = sort_err ? __ldcg(sort_err) : 0u;
    __threadfence_system();
  }
}

struct TableSet {
  GridSlot* table[kMaxLevels];
  uint32_t mask[kMaxLevels];
  uint32_t shift[kMaxLevels];
  int nlevels;
};

// nullptr: the table is full.  Only a table sized SPECULATIVELY (from the previous cloud's cell counts, before
// this cloud's are known) can be: the host sees the real counts a moment later and rebuilds it (build_phase3).
__device__ __forceinline__ GridSlot* slot_insert_or_find(GridSlot* tab, uint32_t mask, uint32_t shift, uint64_t key) {
  uint32_t h = slot_of(key, shift);
  // bounded walk: tables are sized for <= 50 % load (<= 70 % when filled with the previous cloud's sizes, checked by
  // build_phase3 against the real cell counts), where a linear-probing cluster of 2048 slots does not occur
  // (P ~ exp(-0.057 L)); the bound only keeps a speculative fill of a table that turns out too small — which
  // build_phase3 then throws away — from probing the whole table once per insert
  const uint32_t max_probes = mask < 2047u ? mask : 2047u;
  for (uint32_t probes = 0; probes <= max_probes; probes++) {
    unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&tab[h].key), (unsigned long long)kEmptyKey, (unsigned long long)key);
    // the top byte of a live key word collects the occupied-children bits while the table is being built
    if (prev == kEmptyKey || (prev & kKeyMask) == key) return &tab[h];
    h = (h + 1) & mask;
  }
  return nullptr;
}

// blockIdx.y = level: one (point, level) pair per thread, so no thread walks all the levels
// serially (the 1-D version was bound by the CAS chain of the few threads that open a cell at
// every level: 62 us whatever the cloud size).  The thread that opens a cell also sets the cell's bit in
// its parent's occupied-children mask (insert-or-find: whoever comes first creates the parent's slot) —
// this used to be a second kernel over the finished tables (k_child_masks, 13 us on the 500k-point submap).
__global__ void __launch_bounds__(256) k_build_tables(const uint64_t* __restrict__ keys, int n, TableSet ts) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = blockIdx.y;
  if (i >= n) return;
  const uint64_t key = keys[i];
  bool opens = true;  // does point i open a new cell at level l?
  uint64_t prev = 0;
  if (i > 0) {
    prev = keys[i - 1];
    opens = (key >> (3 * l)) != (prev >> (3 * l));
  }
  if (opens) {
    const uint64_t ck = key >> (3 * l);
    GridSlot* s = slot_insert_or_find(ts.table[l], ts.mask[l], ts.shift[l], ck);
    if (s) s->start = (uint32_t)i;
    if (i > 0 && (s = slot_insert_or_find(ts.table[l], ts.mask[l], ts.shift[l], prev >> (3 * l)))) s->end = (uint32_t)i;
    if (l + 1 < ts.nlevels) {
      GridSlot* ps = slot_insert_or_find(ts.table[l + 1], ts.mask[l + 1], ts.shift[l + 1], ck >> 3);
      // the mask lives in the top byte of the 64-bit key word = top byte of its high 32-bit half
      if (ps) atomicOr(reinterpret_cast<unsigned int*>(&ps->key) + 1, 1u << (24 + (int)(ck & 7)));
    }
  }
  if (i == n - 1) {
    GridSlot* s = slot_insert_or_find(ts.table[l], ts.mask[l], ts.shift[l], key >> (3 * l));
    if (s) s->end = (uint32_t)n;
  }
}

// ------------------------------------------------------------------------------------------------
// kNN.  SELF: queries are the grid's own sorted points (thread t <-> sorted point t) and the
// result is stored k-major as sorted positions for k_covariance.  Otherwise queries are float4
// and results go out row-major [m][k] as ORIGINAL indices + d2 (test hook / public rgc_knn).
template <bool SELF>
__global__ void __launch_bounds__(kThreads, 8) k_knn(GridView g, const float4* __restrict__ queries, int m, int k, int spread, int* __restrict__ out_idx,
                                                     float* __restrict__ out_d2) {
  extern __shared__ float heap_smem[];  // [k][kThreads] distances, then [k][kThreads] positions
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  if (gt & (spread - 1)) return;
  const int t = gt / spread;
  if (t >= m) return;
  float4 q = SELF ? reinterpret_cast<const float4*>(g.pts)[t] : queries[t];
  HeapK top;
  top.init(heap_smem + threadIdx.x, reinterpret_cast<int*>(heap_smem + (size_t)k * kThreads) + threadIdx.x, kThreads);
  knn_search(g, q.x, q.y, q.z, k, INFINITY, SELF ? t : -1, top);
  top.sort_ascending(g.pts);
  if (SELF) {
    for (int j = 0; j < k; j++) out_idx[(size_t)j * m + t] = j < top.cnt ? top.id[j * kThreads] : -1;
  } else {
    for (int j = 0; j < k; j++) {
      const bool have = j < top.cnt;
      out_idx[(size_t)t * k + j] = have ? __float_as_int(g.pts[top.id[j * kThreads]].w) : -1;
      if (out_d2) out_d2[(size_t)t * k + j] = have ? top.d[j * kThreads] : INFINITY;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative self-kNN ("tile" kernel): one warp owns 32 Morton-adjacent queries.
//
// An ncu profile of the thread-per-query walk (profiles/) showed ~5 of 32 lanes active: every lane
// walks its own tree and inserts into its own heap at different times.  Here the warp does the
// walk ONCE for its 32 queries and all divergent work is turned into lockstep phases:
//   seeds    the 64 Morton-adjacent points are folded into every lane's heap (lockstep inserts):
//            a tight per-lane bound on the k-th distance before any tree node is touched;
//   gather   warp-uniform depth-first walk; a cell is visited iff ANY lane's ball (its query,
//            radius = its current k-th distance) touches the cell's box (per-lane test + vote);
//            the points of small cells are copied, coalesced, into a shared candidate buffer;
//   consume  every lane scans the buffer (shared-memory broadcast reads, no divergence) and APPENDS
//            the candidates inside its ball to a pending list (no heap work here);
//   fold     when a pending list fills up (and at the end) all lanes fold their pending entries
//            into their heaps together, tightening the balls.
// Exactness is unchanged: a point is dropped only if its cell is outside every ball or its d2
// exceeds the lane's current k-th distance; ties are resolved by (d2, original index) in the heap.
#ifndef RGC_KT_CAND
#define RGC_KT_CAND 128
#endif
#ifndef RGC_KT_STACK
#define RGC_KT_STACK 40
#endif
#ifndef RGC_KT_SEEDS
#define RGC_KT_SEEDS 128
#endif
#ifndef RGC_KT_PEND
#define RGC_KT_PEND 8
#endif
#ifndef RGC_KT_LEAF
#define RGC_KT_LEAF 64
#endif
// defaults from a parameter sweep on the C2 workload (tools/tile_variants.sh; profiles/README.md)
constexpr int KT_WARPS = 4;
constexpr int KT_CAND = RGC_KT_CAND;    // candidate buffer entries per warp (>= KT_SEEDS)
constexpr int KT_STACK = RGC_KT_STACK;  // DFS stack entries per warp
constexpr int KT_SEEDS = RGC_KT_SEEDS;  // Morton-adjacent seed points per warp
constexpr int KT_PEND = RGC_KT_PEND;    // pending slots per lane
constexpr int KT_LEAF = RGC_KT_LEAF;    // cells with <= this many points are gathered whole

struct TileNode {
  uint32_t cx_lvl, cy_mask, cz, start, end;
};
// one tile of a multi-cloud grid (batched registration): 32 consecutive sorted positions starting at
// `first`, all inside cloud [lo, hi) whose cells carry `prefix` (rgc_grid.cuh: CloudRange)
struct TileDesc {
  int first, lo, hi, pad;
  uint64_t prefix;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(KT_WARPS * 32) k_knn_tile(GridView g, int n, int k, int n_seeds, int defer_cands, int* __restrict__ defer_count, int* __restrict__ defer_tiles,
                                                           int* __restrict__ out_idx, const TileDesc* __restrict__ tiles, int ntiles) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long dbg_t0 = clock64();
  unsigned long long dbg_g0 = 0;
  if (g_tile_dbg) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
  int dbg_nodes = 0, dbg_cands = 0, dbg_ins = 0;
  const int cap = k + KT_PEND;
  // per-warp carve-up: candidate buffer | DFS stack | [cap][32] packed heap / pending keys
  const size_t per_warp = sizeof(float4) * KT_CAND + sizeof(TileNode) * KT_STACK + (size_t)cap * 32 * 8;
  unsigned char* base = tile_smem + (size_t)warp * per_warp;
  float4* cand = reinterpret_cast<float4*>(base);
  TileNode* stack = reinterpret_cast<TileNode*>(base + sizeof(float4) * KT_CAND);
  unsigned long long* hk = reinterpret_cast<unsigned long long*>(base + sizeof(float4) * KT_CAND + sizeof(TileNode) * KT_STACK);

  const int tile_id = blockIdx.x * KT_WARPS + warp;
  int first = tile_id * 32, lo = 0, hi = n;
  uint64_t prefix = 0ull;
  if (tiles) {  // multi-cloud grid: the tile list says where each tile lives
    if (tile_id >= ntiles) return;
    const TileDesc td = tiles[tile_id];
    first = td.first;
    lo = td.lo;
    hi = td.hi;
    prefix = td.prefix;
  } else if (first >= n) {
    return;
  }
  const int t = min(first + lane, hi - 1);  // tail lanes shadow the last query (no output)
  const float4 q = reinterpret_cast<const float4*>(g.pts)[t];
  const float4* pts4 = reinterpret_cast<const float4*>(g.pts);

  HeapK64 heap;
  heap.init(hk + lane, 32, k);
  int npend = 0;

  // ---- seeds: the Morton neighbours of the tile go through the same append / fold path as every
  // other candidate (the first k fill the heap, later ones are appended only if they beat the
  // current k-th best), which gives every lane a tight bound before any tree node is touched
  const int ns = min(n_seeds, hi - lo);
  const int s0 = max(lo, min(first - (n_seeds - 32) / 2, hi - ns));
  for (int j = lane; j < ns; j += 32) cand[j] = pts4[s0 + j];
  auto fold = [&]() {
    const int mx = __reduce_max_sync(0xffffffffu, npend);
    for (int e = 0; e < mx; e++)
      if (e < npend) heap.insert(hk[(k + e) * 32 + lane]);
    dbg_ins += mx;
    npend = 0;
  };
  {
    __syncwarp();
    unsigned long long bk = ~0ull;
    for (int j = 0; j < ns; j++) {
      const float4 c = cand[j];
      const unsigned long long key = pack_key(dist2_ref(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w));
      if (key < bk) {
        hk[(k + npend) * 32 + lane] = key;
        npend++;
      }
      if (__any_sync(0xffffffffu, npend == KT_PEND)) {
        fold();
        bk = heap.cnt == k ? hk[lane] : ~0ull;
      }
    }
    fold();
    __syncwarp();
  }

  // ---- geometric cap: the smallest cell around q that holds >= k points bounds the k-th distance
  // by its diagonal.  Lanes whose Morton neighbours lie across a Z-curve jump get a loose bound from
  // the seeds (per-warp stats showed 1 % of the tiles gathering 10-40x the usual candidates because
  // of one such lane); this cap costs a handful of probes and removes that tail.
  float cap2 = INFINITY;
  {
    const int fx = cell_coord(q.x, g.inv_s0, g.bias), fy = cell_coord(q.y, g.inv_s0, g.bias), fz = cell_coord(q.z, g.inv_s0, g.bias);
    const float seed_b = heap.cnt == k ? key_d2(hk[lane]) : INFINITY;
    for (int l = 0; l < g.nlevels; l++) {
      const float edge = g.s0 * pow2f(l) + 2.f * g.margin;
      const float diag2 = 3.f * edge * edge * 1.0001f;
      if (diag2 >= seed_b) break;  // cannot improve on the seed bound any more
      uint32_t s, e, m;
      if (grid_lookup(g, l, fx >> l, fy >> l, fz >> l, s, e, m, prefix >> (3 * l)) && (int)(e - s) >= k) {
        cap2 = diag2;
        break;
      }
    }
  }
  const unsigned long long capk = pack_key(cap2, 0x7fffffff);
  // current k-th best as a key (inclusive bound for appends) and as a distance (for box tests)
  auto bound_key = [&]() { return heap.cnt == k ? hk[lane] : ~0ull; };
  auto bound = [&]() { return fminf(heap.cnt == k ? key_d2(hk[lane]) : INFINITY, cap2); };
  int ncand = 0;
  auto consume = [&]() {
    __syncwarp();
    unsigned long long bk = bound_key();
    for (int j = 0; j < ncand; j++) {
      const float4 c = cand[j];
      const unsigned long long key = pack_key(dist2_ref(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w));
      if (key < bk && key <= capk) {
        hk[(k + npend) * 32 + lane] = key;
        npend++;
      }
      if (__any_sync(0xffffffffu, npend == KT_PEND)) {
        fold();
        bk = bound_key();
      }
    }
    ncand = 0;
    __syncwarp();
  };

  if (hi - lo > ns) {
    // ---- roots: cells covering the union of the balls, at a level where that is <= 4 cells per axis
    const float b0 = bound();
    const float r = b0 < INFINITY ? sqrtf(b0) * 1.00001f + 2.f * g.margin : INFINITY;
    const float lox = warp_min(q.x - r), loy = warp_min(q.y - r), loz = warp_min(q.z - r);
    const float hix = warp_max(q.x + r), hiy = warp_max(q.y + r), hiz = warp_max(q.z + r);
    const float ext = fmaxf(fmaxf(hix - lox, hiy - loy), hiz - loz);
    const int top_level = g.nlevels - 1;
    const int lb = root_level(g.s0, 0.5f * ext, top_level);
    const float inv_cs = g.inv_s0 * pow2f(-lb);
    const int ncell = 1 << (g.nbits - lb);
    int rlo[3], rhi[3];
    {
      const float l3[3] = {lox, loy, loz}, h3[3] = {hix, hiy, hiz}, o3[3] = {g.ox, g.oy, g.oz};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const float tl = (l3[a] - o3[a]) * inv_cs, th = (h3[a] - o3[a]) * inv_cs;
        const int il = tl < 0.f ? 0 : (tl >= (float)ncell ? ncell : (int)tl);
        const int ih = th < 0.f ? -1 : (th >= (float)ncell ? ncell - 1 : (int)th);
        rlo[a] = il;
        rhi[a] = ih < ncell - 1 ? ih : ncell - 1;
        if (ih < 0 || il >= ncell) rhi[a] = rlo[a] - 1;
      }
    }
    // root cells are resolved 32 at a time: lane r tests / looks up root r, hits are pushed
    const int nx = rhi[0] - rlo[0] + 1, ny = rhi[1] - rlo[1] + 1, nz = rhi[2] - rlo[2] + 1;
    const int nroots = (nx > 0 && ny > 0 && nz > 0) ? nx * ny * nz : 0;
    int sp = 0;
    for (int r0 = 0; r0 < nroots; r0 += 32) {
      const int ri = r0 + lane;
      uint32_t s = 0, e = 0, m = 0;
      int rx = 0, ry = 0, rz = 0;
      bool have = false;
      if (ri < nroots) {
        rx = rlo[0] + ri % nx;
        ry = rlo[1] + (ri / nx) % ny;
        rz = rlo[2] + ri / (nx * ny);
        have = grid_lookup(g, lb, rx, ry, rz, s, e, m, prefix >> (3 * lb));
      }
      uint32_t hv = __ballot_sync(0xffffffffu, have);
      while (hv) {
        const int src_lane = __ffs(hv) - 1;
        hv &= hv - 1;
        // does any lane's ball touch this root?  (every lane tests its own query)
        const int bx = __shfl_sync(0xffffffffu, rx, src_lane), by = __shfl_sync(0xffffffffu, ry, src_lane), bz = __shfl_sync(0xffffffffu, rz, src_lane);
        if (!__any_sync(0xffffffffu, box_dist2(g, lb, bx, by, bz, q.x, q.y, q.z) <= bound())) continue;
        if (sp < KT_STACK) {
          if (lane == src_lane) stack[sp] = TileNode{(uint32_t)rx | ((uint32_t)lb << 24), (uint32_t)ry | (m << 24), (uint32_t)rz, s, e};
          sp++;
        }
      }
      __syncwarp();
      // ---- depth-first walk of everything pushed so far
      while (sp > 0) {
        if (dbg_cands > defer_cands) {
          // this tile's queries do not share a neighbourhood (sparse region: the union of the 32
          // balls keeps growing); hand it to k_knn_warp, one warp per query, instead of becoming
          // the tail of this launch
          if (lane == 0) {
            defer_tiles[atomicAdd(defer_count, 1)] = tile_id;
            if (g_tile_dbg) {
              long long* o = g_tile_dbg + (size_t)tile_id * 4;
              o[0] = clock64() - dbg_t0;
              o[1] = dbg_nodes;
              o[2] = (long long)dbg_cands | (1ll << 40);
              o[3] = (long long)dbg_ins | ((long long)((dbg_g0 >> 6) & 0x3fffffff) << 32);
            }
          }
          return;
        }
        const TileNode nd = stack[--sp];
        __syncwarp();
        dbg_nodes++;
        const int l = (int)(nd.cx_lvl >> 24);
        const int cx = (int)(nd.cx_lvl & 0xffffffu), cy = (int)(nd.cy_mask & 0xffffffu), cz = (int)nd.cz;
        const uint32_t cm = nd.cy_mask >> 24;
        if (!__any_sync(0xffffffffu, box_dist2(g, l, cx, cy, cz, q.x, q.y, q.z) <= bound())) continue;
        if (l == 0 || nd.end - nd.start <= (uint32_t)KT_LEAF || sp + 8 > KT_STACK) {
          // gather the cell's points (minus the seeds, already folded)
          for (uint32_t p0 = nd.start; p0 < nd.end; p0 += 32) {
            const uint32_t p = p0 + lane;
            const bool take = p < nd.end && (uint32_t)((int)p - s0) >= (uint32_t)ns;
            const uint32_t bal = __ballot_sync(0xffffffffu, take);
            if (take) cand[ncand + __popc(bal & ((1u << lane) - 1u))] = pts4[p];
            ncand += __popc(bal);
            dbg_cands += __popc(bal);
            if (ncand > KT_CAND - 32) consume();
          }
          continue;
        }
        // expand: lane c < 8 resolves child c (mask bit -> hash lookup); the others wait
        const uint64_t pkey = (morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz) | (prefix >> (3 * l))) << 3;
        uint32_t cs = 0, ce = 0, cmk = 0;
        bool hc = false;
        if (lane < 8 && ((cm >> lane) & 1u)) hc = grid_lookup_key(g, l - 1, pkey | (uint64_t)lane, cs, ce, cmk);
        const uint32_t hvc = __ballot_sync(0xffffffffu, hc);
        for (int c = 0; c < 8; c++) {
          if (!((hvc >> c) & 1u)) continue;
          const int ccx = 2 * cx + (c & 1), ccy = 2 * cy + ((c >> 1) & 1), ccz = 2 * cz + ((c >> 2) & 1);
          if (!__any_sync(0xffffffffu, box_dist2(g, l - 1, ccx, ccy, ccz, q.x, q.y, q.z) <= bound())) continue;
          if (lane == c) stack[sp] = TileNode{(uint32_t)ccx | ((uint32_t)(l - 1) << 24), (uint32_t)ccy | (cmk << 24), (uint32_t)ccz, cs, ce};
          sp++;
        }
        __syncwarp();
      }
    }
    consume();
  }
  fold();
  heap.sort_ascending();
  if (first + lane < hi)
    for (int j = 0; j < k; j++) out_idx[(size_t)j * n + t] = j < heap.cnt ? __ldg(&g.inv[(unsigned)(hk[j * 32 + lane] & 0xffffffffull)]) : -1;
  if (g_tile_dbg && lane == 0) {
    long long* o = g_tile_dbg + (size_t)tile_id * 4;
    // level of the smallest cell holding the whole tile, and the launch-relative start time
    const float4 pa = pts4[first], pb = pts4[min(first + 31, hi - 1)];
    const uint64_t ma = morton3(cell_coord(pa.x, g.inv_s0, g.bias), cell_coord(pa.y, g.inv_s0, g.bias), cell_coord(pa.z, g.inv_s0, g.bias));
    const uint64_t mb = morton3(cell_coord(pb.x, g.inv_s0, g.bias), cell_coord(pb.y, g.inv_s0, g.bias), cell_coord(pb.z, g.inv_s0, g.bias));
    const int lca = ma == mb ? 0 : (63 - __clzll((long long)(ma ^ mb))) / 3 + 1;
    o[0] = clock64() - dbg_t0;
    o[1] = (long long)dbg_nodes | ((long long)lca << 40);
    o[2] = dbg_cands;
    o[3] = (long long)dbg_ins | ((long long)((dbg_g0 >> 6) & 0x3fffffff) << 32);
  }
}

// ------------------------------------------------------------------------------------------------
// One warp per query, for the tiles k_knn_tile deferred (k <= 32).  The k best keys live in
// registers, lane j holding the j-th smallest; candidates are tested 32 at a time and the few that
// pass are inserted with a ballot (rank) and one shuffle (shift).  Tree nodes are expanded with
// lanes 0-7 probing the eight children in parallel.  Same keys, same distance arithmetic, same
// conservative box test as the tile kernel, hence the same neighbour lists.
constexpr int KW_WARPS = 8;
constexpr int KW_STACK = 64;
constexpr int KW_LEAF = 64;
#ifndef RGC_KW_SEEDS
#define RGC_KW_SEEDS 64
#endif
constexpr int KW_SEEDS = RGC_KW_SEEDS;
#ifndef RGC_KW_MERGE_MIN
#define RGC_KW_MERGE_MIN 6
#endif
constexpr int KW_MERGE_MIN = RGC_KW_MERGE_MIN;  // a batch with more qualifying candidates than this is merged by a sorting network

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  return ((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(v >> 32), src) << 32) | __shfl_sync(0xffffffffu, (unsigned)v, src);
}
__device__ __forceinline__ unsigned long long shfl64_xor(unsigned long long v, int m) {
  return ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
}
__device__ __forceinline__ unsigned long long shfl64_up1(unsigned long long v) {
  return ((unsigned long long)__shfl_up_sync(0xffffffffu, (unsigned)(v >> 32), 1) << 32) | __shfl_up_sync(0xffffffffu, (unsigned)v, 1);
}

// Two ways of naming the queries: `defer_tiles` (tiles of 32 consecutive sorted positions handed over by
// k_knn_tile; results k-major with stride n at the query's own position) or, with `qlist` set, an
// explicit list of sorted positions (on-demand target covariances: the correspondences of one
// linearize; results k-major with stride `out_stride` at the LIST index).
// Multi-cloud grids: `tiles` (deferred mode) or `cloud_off` (list mode: n_clouds + 1 sorted-position
// offsets, the query's cloud is found by bisection) confine every search to the query's own cloud.
#ifndef RGC_KW_MINB
#define RGC_KW_MINB 4
#endif
__global__ void __launch_bounds__(KW_WARPS * 32, RGC_KW_MINB) k_knn_warp(GridView g, int n, int k, const int* __restrict__ defer_count, const int* __restrict__ defer_tiles,
                                                           const int* __restrict__ qlist, int out_stride, int* __restrict__ out_idx,
                                                           const TileDesc* __restrict__ tiles, const int* __restrict__ cloud_off, int n_clouds) {
  pdl_enter();
  __shared__ TileNode stacks[KW_WARPS][KW_STACK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TileNode* stack = stacks[warp];
  const float4* pts4 = reinterpret_cast<const float4*>(g.pts);
  const int nq = qlist ? *defer_count : (defer_count ? *defer_count * 32 : n);  // no list at all: every point of the cloud
  for (int w = blockIdx.x * KW_WARPS + warp; w < nq; w += gridDim.x * KW_WARPS) {
    int t, lo = 0, hi = n;
    uint64_t prefix = 0ull;
    if (qlist) {
      t = qlist[w];
      if (cloud_off) {
        int a = 0, b = n_clouds;  // cloud_off[a] <= t < cloud_off[b]
        while (b - a > 1) {
          const int mid = (a + b) >> 1;
          if (cloud_off[mid] <= t) a = mid; else b = mid;
        }
        lo = cloud_off[a];
        hi = cloud_off[a + 1];
        prefix = (uint64_t)a << (3 * g.nbits);
      }
    } else if (tiles) {
      const TileDesc td = tiles[defer_tiles[w >> 5]];
      t = td.first + (w & 31);
      lo = td.lo;
      hi = td.hi;
      prefix = td.prefix;
    } else if (defer_tiles) {
      t = defer_tiles[w >> 5] * 32 + (w & 31);
    } else {
      t = w;
    }
    if (t >= hi) continue;
    const float4 q = pts4[t];
    unsigned long long mine = ~0ull;  // lane j: j-th smallest key so far (~0 = empty)
    unsigned long long capk = ~0ull;
    float cap2 = INFINITY;
    auto offer = [&](unsigned long long ck, bool valid) {
      unsigned long long wk = shfl64(mine, k - 1);
      unsigned acc = __ballot_sync(0xffffffffu, valid && ck < wk && ck <= capk);
      if (__popc(acc) > KW_MERGE_MIN) {
        // many of the 32 candidates qualify (the seeds of every query; whole batches in sparse regions, whose queries
        // were the tail of the launch: ~150 dependent cycles per one-at-a-time insertion): sort the batch with a
        // bitonic network, keep the 32 smallest of (list, batch) — min(list[i], batch[31 - i]) is a bitonic sequence
        // of exactly those — and merge: 20 compare-exchange steps whatever the number of insertions
        unsigned long long b = ((acc >> lane) & 1u) ? ck : ~0ull;
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
          for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const unsigned long long o = shfl64_xor(b, stride);
            const bool keep_min = ((lane & stride) == 0) == ((lane & size) == 0);
            b = keep_min ? (o < b ? o : b) : (o > b ? o : b);
          }
        const unsigned long long r = shfl64(b, 31 - lane);
        unsigned long long m = r < mine ? r : mine;
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) {
          const unsigned long long o = shfl64_xor(m, stride);
          m = ((lane & stride) == 0) ? (o < m ? o : m) : (o > m ? o : m);
        }
        mine = m;
        return;
      }
      while (acc) {
        const int src = __ffs(acc) - 1;
        const unsigned long long nk = shfl64(ck, src);
        const int pos = __popc(__ballot_sync(0xffffffffu, mine < nk));
        const unsigned long long up = shfl64_up1(mine);
        if (lane == pos) mine = nk;
        else if (lane > pos) mine = up;
        wk = shfl64(mine, k - 1);
        acc &= acc - 1;
        acc &= __ballot_sync(0xffffffffu, ck < wk);
      }
    };
    auto bound = [&]() {
      const unsigned long long wk = shfl64(mine, k - 1);
      return fminf(wk == ~0ull ? INFINITY : key_d2(wk), cap2);
    };
    // seeds: Morton neighbours of the query
    const int ns = min(KW_SEEDS, hi - lo);
    const int s0 = max(lo, min(t - KW_SEEDS / 2, hi - ns));
    for (int j0 = 0; j0 < ns; j0 += 32) {
      const bool take = j0 + lane < ns;
      unsigned long long key = ~0ull;
      if (take) {
        const float4 c = pts4[s0 + j0 + lane];
        key = pack_key(dist2_ref(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w));
      }
      offer(key, take);
    }
    if (hi - lo > ns) {
      // geometric cap (see k_knn_tile): lane l probes the level-l cell of the query
      {
        const int fx = cell_coord(q.x, g.inv_s0, g.bias), fy = cell_coord(q.y, g.inv_s0, g.bias), fz = cell_coord(q.z, g.inv_s0, g.bias);
        bool enough = false;
        if (lane < g.nlevels) {
          uint32_t s, e, m;
          enough = grid_lookup(g, lane, fx >> lane, fy >> lane, fz >> lane, s, e, m, prefix >> (3 * lane)) && (int)(e - s) >= k;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, enough);
        if (hit) {
          const int l = __ffs(hit) - 1;
          const float edge = g.s0 * pow2f(l) + 2.f * g.margin;
          cap2 = 3.f * edge * edge * 1.0001f;
          capk = pack_key(cap2, 0x7fffffff);
        }
      }
      const float b0 = bound();
      const float r = b0 < INFINITY ? sqrtf(b0) * 1.00001f + 2.f * g.margin : INFINITY;
      const int top_level = g.nlevels - 1;
      const int lb = root_level(g.s0, r, top_level);
      const float inv_cs = g.inv_s0 * pow2f(-lb);
      const int ncell = 1 << (g.nbits - lb);
      int rlo[3], rhi[3];
      {
        const float q3[3] = {q.x, q.y, q.z}, o3[3] = {g.ox, g.oy, g.oz};
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const float tl = (q3[a] - r - o3[a]) * inv_cs, th = (q3[a] + r - o3[a]) * inv_cs;
          const int il = !(tl >= 0.f) ? 0 : (tl >= (float)ncell ? ncell : (int)tl);
          const int ih = !(th >= 0.f) ? (th != th || r == INFINITY ? ncell - 1 : -1) : (th >= (float)ncell ? ncell - 1 : (int)th);
          rlo[a] = il;
          rhi[a] = ih < ncell - 1 ? ih : ncell - 1;
          if (ih < 0 || il >= ncell) rhi[a] = rlo[a] - 1;
        }
      }
      const int nx = rhi[0] - rlo[0] + 1, ny = rhi[1] - rlo[1] + 1, nz = rhi[2] - rlo[2] + 1;
      const int nroots = (nx > 0 && ny > 0 && nz > 0) ? nx * ny * nz : 0;
      int sp = 0;
      for (int r0 = 0; r0 < nroots; r0 += 32) {
        const int ri = r0 + lane;
        uint32_t s = 0, e = 0, m = 0;
        int rx = 0, ry = 0, rz = 0;
        bool have = false;
        if (ri < nroots) {
          rx = rlo[0] + ri % nx;
          ry = rlo[1] + (ri / nx) % ny;
          rz = rlo[2] + ri / (nx * ny);
          have = grid_lookup(g, lb, rx, ry, rz, s, e, m, prefix >> (3 * lb));
        }
        const float bnd = bound();
        have = have && box_dist2(g, lb, rx, ry, rz, q.x, q.y, q.z) <= bnd;
        const unsigned hv = __ballot_sync(0xffffffffu, have);
        const int room = KW_STACK - sp;
        const int slot = __popc(hv & ((1u << lane) - 1u));
        if (have && slot < room) stack[sp + slot] = TileNode{(uint32_t)rx | ((uint32_t)lb << 24), (uint32_t)ry | (m << 24), (uint32_t)rz, s, e};
        const int pushed = min(__popc(hv), room);
        // roots that do not fit on the stack (cannot happen for nroots <= 27 unless a previous
        // batch left entries) are scanned directly
        unsigned over = hv;
        for (int i = 0; i < pushed; i++) over &= over - 1;
        sp += pushed;
        __syncwarp();
        while (over) {
          const int src = __ffs(over) - 1;
          over &= over - 1;
          const uint32_t os = __shfl_sync(0xffffffffu, s, src), oe = __shfl_sync(0xffffffffu, e, src);
          for (uint32_t p0 = os; p0 < oe; p0 += 32) {
            const uint32_t p = p0 + lane;
            const bool take = p < oe && (uint32_t)((int)p - s0) >= (uint32_t)ns;
            unsigned long long key = ~0ull;
            if (take) {
              const float4 c = pts4[p];
              key = pack_key(dist2_ref(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w));
            }
            offer(key, take);
          }
        }
        // Depth-first walk, up to four nodes per step: lane group g = lane / 8 takes the g-th node from the top of
        // the stack.  Leaf nodes of the step are scanned together (32 candidates per round over the
        // concatenation of their ranges) and the inner nodes are expanded together (lane 8 g + c resolves
        // and tests child c of group g's node: 32 table probes in flight instead of 8).  One query is a
        // chain of dependent memory accesses (a short on-demand list ran 17 us however few queries it held);
        // stepping four nodes at a time cuts the number of links.  Any visiting order gives the same exact
        // neighbour list: a node is only ever skipped against a bound that is already a valid upper bound.
        while (sp > 0) {
          int nb = min(sp, 4);
          while (nb > 1 && (sp - nb) + 8 * nb > KW_STACK) nb--;
          const int grp = lane >> 3, sub = lane & 7;
          const bool in = grp < nb;
          const TileNode nd = stack[sp - 1 - (in ? grp : 0)];
          sp -= nb;
          __syncwarp();
          const int l = (int)(nd.cx_lvl >> 24);
          const int cx = (int)(nd.cx_lvl & 0xffffffu), cy = (int)(nd.cy_mask & 0xffffffu), cz = (int)nd.cz;
          const uint32_t cm = nd.cy_mask >> 24;
          const float bd = bound();
          const bool live = in && box_dist2(g, l, cx, cy, cz, q.x, q.y, q.z) <= bd;
          const bool leaf = live && (l == 0 || nd.end - nd.start <= (uint32_t)KW_LEAF || (nb == 1 && sp + 8 > KW_STACK));
          // ---- leaves of this step
          const uint32_t lsz = leaf ? nd.end - nd.start : 0u;
          const uint32_t z0 = __shfl_sync(0xffffffffu, lsz, 0), z1 = __shfl_sync(0xffffffffu, lsz, 8), z2 = __shfl_sync(0xffffffffu, lsz, 16),
                         z3 = __shfl_sync(0xffffffffu, lsz, 24);
          const uint32_t total = z0 + z1 + z2 + z3;
          if (total) {
            const uint32_t b0s = __shfl_sync(0xffffffffu, nd.start, 0), b1s = __shfl_sync(0xffffffffu, nd.start, 8),
                           b2s = __shfl_sync(0xffffffffu, nd.start, 16), b3s = __shfl_sync(0xffffffffu, nd.start, 24);
            for (uint32_t e0 = 0; e0 < total; e0 += 32) {
              const uint32_t e = e0 + lane;
              uint32_t p;
              if (e < z0) p = b0s + e;
              else if (e < z0 + z1) p = b1s + (e - z0);
              else if (e < z0 + z1 + z2) p = b2s + (e - z0 - z1);
              else p = b3s + (e - z0 - z1 - z2);
              const bool take = e < total && (uint32_t)((int)p - s0) >= (uint32_t)ns;
              unsigned long long key = ~0ull;
              if (take) {
                const float4 c = pts4[p];
                key = pack_key(dist2_ref(q.x, q.y, q.z, c.x, c.y, c.z), __float_as_int(c.w));
              }
              offer(key, take);
            }
          }
          // ---- inner nodes of this step: lane 8 g + c resolves and tests child c of group g's node; the survivors
          // of a group are pushed farthest first, the groups deepest first, so the nearest child of the node that
          // was on top of the stack is popped next
          const bool inner = live && !leaf;
          if (!__any_sync(0xffffffffu, inner)) continue;
          const float bd2 = total ? bound() : bd;
          uint32_t cs = 0, ce = 0, cmk = 0;
          bool hc = false;
          float dc = INFINITY;
          const int ccx = 2 * cx + (sub & 1), ccy = 2 * cy + ((sub >> 1) & 1), ccz = 2 * cz + ((sub >> 2) & 1);
          if (inner && ((cm >> sub) & 1u)) {
            const uint64_t pkey = (morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz) | (prefix >> (3 * l))) << 3;
            hc = grid_lookup_key(g, l - 1, pkey | (uint64_t)sub, cs, ce, cmk);
            if (hc) {
              dc = box_dist2(g, l - 1, ccx, ccy, ccz, q.x, q.y, q.z);
              hc = dc <= bd2;
            }
          }
          const unsigned hvc = __ballot_sync(0xffffffffu, hc);
          const unsigned mine_grp = (hvc >> (8 * grp)) & 0xffu;
          int rank = 0;
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float dother = __shfl_sync(0xffffffffu, dc, c, 8);
            if (((mine_grp >> c) & 1u) && (dother > dc || (dother == dc && c < sub))) rank++;
          }
          // groups above this one in the stack order (pushed earlier): those with a HIGHER group index
          const int before = grp < 3 ? __popc(hvc >> (8 * (grp + 1))) : 0;
          if (hc) stack[sp + before + rank] = TileNode{(uint32_t)ccx | ((uint32_t)(l - 1) << 24), (uint32_t)ccy | (cmk << 24), (uint32_t)ccz, cs, ce};
          sp += __popc(hvc);
          __syncwarp();
        }
      }
    }
    if (lane < k) out_idx[qlist ? (size_t)lane * out_stride + w : (size_t)lane * n + t] = mine != ~0ull ? __ldg(&g.inv[(unsigned)(mine & 0xffffffffull)]) : -1;
    __syncwarp();
  }
}

// covariance of sorted point t from its k neighbour positions (k-major), regularised; 6 doubles out.
// All k index loads are issued first, then all k point gathers (fully unrolled, KCAP is a compile
// time bound on k): the first version walked the neighbours in a serial loop of dependent
// index -> point loads and was latency-bound at 17 % of HBM peak however cheap the fp64 part became.
#ifndef RGC_COV_MINB  // blocks per SM the register allocation aims at (the staged variant is limited to 5 by its 40 KB of shared memory)
#define RGC_COV_MINB 6
#endif
#ifndef RGC_COV_BATCH
#define RGC_COV_BATCH 5
#endif
// FULL: k == KCAP and the cloud has at least k points, so every neighbour slot is valid (no predicates).
// `qlist` / `qcount` (nullable): compute only the listed sorted positions (on-demand mode); thread e then
// reads its neighbours at nbr[j * n + e] (n = the list's stride) and writes cov[qlist[e]].  Same code,
// same neighbour order, hence bit-identical values to the whole-cloud launch.
template <int KCAP, bool FULL>
__global__ void __launch_bounds__(kThreads, RGC_COV_MINB) k_covariance(const float4* __restrict__ pts, const int* __restrict__ nbr, int n, int k, int method,
                                                            double* __restrict__ cov, const int* __restrict__ qlist, const int* __restrict__ qcount) {
  pdl_enter();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (qlist ? *qcount : n)) return;
  // all k index loads first (coalesced, k-major), then the point gathers in batches of kBatch:
  // a float4 gather holds four registers while in flight, 20 at once cost 128 registers / thread
  // and a third of the occupancy
  constexpr int kBatch = RGC_COV_BATCH;
  int id[KCAP];
#pragma unroll
  for (int j = 0; j < KCAP; j++) id[j] = (FULL || j < k) ? __ldg(&nbr[(size_t)j * n + t]) : -1;
  // same arithmetic as covariance_from_points (rgc_math.cuh): moments about neighbour 0, then re-centre
  int found = 0;
  Sym3 c = {0, 0, 0, 0, 0, 0};
  double ox = 0.0, oy = 0.0, oz = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
#if RGC_ASYNC_STAGE
  // all k neighbour points of this thread are copied global -> shared at once (cp.async: no registers in flight; the
  // register version had to gather in batches of five, four dependent round trips per point), neighbour-major so
  // the 128-bit reads back are conflict-free.  KCAP = 20: 40 KB per block.
  extern __shared__ __align__(16) float4 nb_stage[];  // [KCAP][kThreads]
#pragma unroll
  for (int j = 0; j < KCAP; j++)
    if (FULL || id[j] >= 0) cp_async16(&nb_stage[j * kThreads + threadIdx.x], &pts[id[j]], true);
  cp_async_commit();
  cp_async_wait<0>();
#endif
#pragma unroll
  for (int j0 = 0; j0 < KCAP; j0 += kBatch) {
    float px[kBatch], py[kBatch], pz[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int j = j0 + u;
      px[u] = py[u] = pz[u] = 0.f;
      if (j < KCAP && (FULL || id[j] >= 0)) {
#if RGC_ASYNC_STAGE
        const float4 p = nb_stage[j * kThreads + threadIdx.x];
#else
        const float4 p = __ldg(&pts[id[j]]);
#endif
        px[u] = p.x;
        py[u] = p.y;
        pz[u] = p.z;
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int j = j0 + u;
      if (j >= KCAP || (!FULL && id[j] < 0)) continue;  // valid entries are a prefix
      found++;
      if (j == 0) {
        ox = f2d(px[0]);
        oy = f2d(py[0]);
        oz = f2d(pz[0]);
        continue;
      }
      const double dx = f2d(px[u]) - ox, dy = f2d(py[u]) - oy, dz = f2d(pz[u]) - oz;
      sx += dx;
      sy += dy;
      sz += dz;
      c.xx += dx * dx;
      c.xy += dx * dy;
      c.xz += dx * dz;
      c.yy += dy * dy;
      c.yz += dy * dz;
      c.zz += dz * dz;
    }
  }
  if (found > 0) {
    const int missing = k - found;
    if (missing > 0) {
      const double m = (double)missing;
      sx -= m * ox;
      sy -= m * oy;
      sz -= m * oz;
      c.xx += m * ox * ox;
      c.xy += m * ox * oy;
      c.xz += m * ox * oz;
      c.yy += m * oy * oy;
      c.yz += m * oy * oz;
      c.zz += m * oz * oz;
    }
    const double ik = 1.0 / (double)k;
    const double mx = sx * ik, my = sy * ik, mz = sz * ik;
    c.xx = c.xx * ik - mx * mx;
    c.xy = c.xy * ik - mx * my;
    c.xz = c.xz * ik - mx * mz;
    c.yy = c.yy * ik - my * my;
    c.yz = c.yz * ik - my * mz;
    c.zz = c.zz * ik - mz * mz;
  }
  const Sym3 r = regularize_cov(c, method);
  double2* o = reinterpret_cast<double2*>(cov + (size_t)(qlist ? qlist[t] : t) * 6);
  o[0] = make_double2(r.xx, r.xy);
  o[1] = make_double2(r.xz, r.yy);
  o[2] = make_double2(r.yz, r.zz);
}

// ---- asynchronous global -> shared copies (cp.async, LDGSTS): 16 bytes per instruction, no registers held while the
// load is in flight.  The streaming kernels of the path (covariance, linearize, compute_error) were latency-bound on
// their gathers (ncu: long-scoreboard stalls 6-7 of ~10 warp-cycles per issue): a thread could only keep in
// flight what it had registers for.  Staging through shared memory lets every thread issue ALL the gathers of
// its next point (20 neighbours / the 128 bytes of a correspondence) at once.  Each thread reads back only the
// slots it filled itself, so cp.async.wait_group is the only synchronisation needed.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool keep_l1) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (keep_l1)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
  else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// MEASURED, NOT THE DEFAULT (round 2, 8 M / 2 M points): staged k_covariance 0.597 ms vs 0.447 ms with register
// gathers in batches of five, k_linearize 0.109 vs 0.094 ms, k_compute_error 0.059 vs 0.045 ms — the extra
// shared-memory round trip and the occupancy lost to the 40 KB / 33 KB of staging cost more than the deeper
// gather queue wins; at 5 or 6 blocks per SM (register cap 96 / 80, spills) k_linearize fell to 0.133 / 0.18 ms.
// The code stays behind RGC_ASYNC_STAGE=1 for A/B runs (profiles/README.md).
#ifndef RGC_ASYNC_STAGE
#define RGC_ASYNC_STAGE 0
#endif

__device__ __forceinline__ Sym3 load_sym3(const double* __restrict__ base, size_t i) {
  const double2* p = reinterpret_cast<const double2*>(base + i * 6);
  double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  return Sym3{a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ void store_sym3(double* __restrict__ base, size_t i, const Sym3& s) {
  double2* p = reinterpret_cast<double2*>(base + i * 6);
  p[0] = make_double2(s.xx, s.xy);
  p[1] = make_double2(s.xz, s.yy);
  p[2] = make_double2(s.yz, s.zz);
}

// any k (32 < k <= 128, whole clouds only): one thread per point walking its neighbour list — the plain
// covariance_from_points of rgc_math.cuh.  The unrolled kernel above keeps the k indices in registers.
__global__ void __launch_bounds__(kThreads) k_covariance_any(const float4* __restrict__ pts, const int* __restrict__ nbr, int n_stride, int n, int k, int method,
                                                             double* __restrict__ cov) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int found = 0;
  while (found < k && __ldg(&nbr[(size_t)found * n_stride + t]) >= 0) found++;  // valid entries are a prefix
  const Sym3 c = covariance_from_points(found, k, [&](int j) {
    const float4 p = __ldg(&pts[__ldg(&nbr[(size_t)j * n_stride + t])]);
    return F4{p.x, p.y, p.z, p.w};
  });
  store_sym3(cov, t, regularize_cov(c, method));
}


// ------------------------------------------------------------------------------------------------
// Deterministic grid reduction of NV doubles per thread: warp shuffle -> smem -> per-block partial
// -> the last block to finish sums the partials in block order and writes `result` (which may be
// mapped pinned host memory).  The order is fixed by the launch shape, so results are
// bit-reproducible run to run (the reference's OpenMP sums are not).
// `done`: optional sequence word in mapped pinned host memory, written after the results are visible
// system-wide, so the host can poll it instead of paying a cudaStreamSynchronize per LM step.
struct DoneFlag {
  unsigned long long* flag;
  unsigned long long value;
};
// `block_id` / `n_blocks`: the position of this block in the (possibly virtual) grid whose partials are
// summed together — the whole launch for the single-registration kernels, one pair's slice of the launch
// for the batched ones (rgc_batch.cuh), which therefore add in exactly the same order.
template <int NV>
__device__ __forceinline__ void grid_reduce_at(double* v, double* __restrict__ partials, unsigned block_id, unsigned n_blocks, unsigned int* __restrict__ ticket,
                                               double* __restrict__ result, DoneFlag done = DoneFlag{nullptr, 0ull}, int* __restrict__ zero_me = nullptr) {
  __shared__ double sm[kThreads / 32][NV];
  __shared__ bool is_last;
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
    double s = sm[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < kThreads / 32; w++) s += sm[w][threadIdx.x];
    partials[(size_t)block_id * NV + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == n_blocks - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    // final pass, all threads: thread t adds the partials of blocks t, t + kThreads, ... (all NV values of a
    // block at a time: short chains of independent loads), then the same warp tree / warp-order sum as above
    // combines the kThreads per-thread sums — a fixed order for a given n_blocks, whichever block runs last.
    // (The first version gave each value to 4 threads that walked n_blocks / 4 partials one dependent add
    // after the other: ~7 us of a 13 us compute_error on one sweep, ~20 us of 110 us at 2 M points.)
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = 0.0;
    for (unsigned b = threadIdx.x; b < n_blocks; b += kThreads) {
      const double* pb = partials + (size_t)b * NV;
#pragma unroll
      for (int j = 0; j < NV; j++) v[j] += __ldcg(pb + j);
    }
    __syncthreads();  // sm is reused
#pragma unroll
    for (int j = 0; j < NV; j++) {
      double x = v[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) sm[warp][j] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
      double tot = sm[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < kThreads / 32; w++) tot += sm[w][threadIdx.x];
      result[threadIdx.x] = tot;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      *ticket = 0;
      if (zero_me) *zero_me = 0;
      if (done.flag) *reinterpret_cast<volatile unsigned long long*>(done.flag) = done.value;
    }
    __threadfence_system();
  }
}

template <int NV>
__device__ __forceinline__ void grid_reduce(double* v, double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ result,
                                            DoneFlag done = DoneFlag{nullptr, 0ull}, int* __restrict__ zero_me = nullptr) {
  grid_reduce_at<NV>(v, partials, blockIdx.x, gridDim.x, ticket, result, done, zero_me);
}

// grid size of the reduction kernels: enough blocks to fill the machine, few enough that the
// final pass over the per-block partials stays short (persistent grid-stride threads)
__host__ __device__ inline int reduce_grid(int n) {
  const int blocks = (n + kThreads - 1) / kThreads;
  return blocks < 148 * 8 ? blocks : 148 * 8;
}

struct RtF {
  float m[12];
};

// ownership slab of a sharded target: a query is this rank's iff lo <= q[axis] < hi (axis < 0: all)
struct Slab {
  int axis;
  float lo, hi;
};
__device__ __forceinline__ bool slab_owns(const Slab& s, float qx, float qy, float qz) {
  if (s.axis < 0) return true;
  const float v = s.axis == 0 ? qx : (s.axis == 1 ? qy : qz);
  return v >= s.lo && v < s.hi;
}

constexpr int kLinN = kAccN + 1;  // + inlier count

// update_correspondences, search part (fast_gicp_impl.hpp:115-137): float transform of every source
// point, exact 1-NN in the target (pruned at thr2), thresholded.  SM/latency-bound tree walk; small
// clouds use sparse warps.  Stores the sorted target position (or -1) and d2.
// `use_hint`: corr[] still holds the previous iteration's correspondences; the old neighbour is an
// excellent first candidate (the pose moves little between LM iterations), so the search starts
// with the tightest possible ball and skips the climb.  Any real point is a valid bound: the
// result is the same exact nearest neighbour.
// `need_state` (nullable): on-demand target covariances.  A correspondence whose target point has no
// covariance yet claims it (0 -> 1) and appends its sorted position to `need_list`; the kNN +
// covariance kernels that follow compute exactly those (linearize reads C_B only at correspondences,
// fast_gicp_impl.hpp:139-146, so the values it sees are the ones the eager pass would have produced).
#ifndef RGC_CORR_MINB
#define RGC_CORR_MINB 8
#endif
#ifndef RGC_GROUP_SEARCH  // 0: the sparse-warp searches walk alone (one lane per query), as before round 2
#define RGC_GROUP_SEARCH 1
#endif
// exact 1-NN of one query by the `spread` lanes that share it (spread 4 / 8: nn1_search_group, every lane of the
// group must call; spread 1 / 2: only the group's first lane gets here and walks alone)
__device__ __forceinline__ void nn1_spread(const GridView& tgt, float qx, float qy, float qz, float max_d2, int hint, int spread, Best1& top) {
  const int lane = threadIdx.x & 31;
  if (spread == 4)
    nn1_search_group<4>(tgt, qx, qy, qz, max_d2, hint, lane & 3, 0xfu << (lane & ~3), top);
  else if (spread == 8)
    nn1_search_group<8>(tgt, qx, qy, qz, max_d2, hint, lane & 7, 0xffu << (lane & ~7), top);
  else
    knn_search(tgt, qx, qy, qz, 1, max_d2, hint, top);
}
__device__ __forceinline__ bool group_search(int spread) { return (spread == 4 || spread == 8) && !g_tile_dbg && RGC_GROUP_SEARCH; }

__device__ __forceinline__ void correspond_query(const GridView& tgt, const float4* __restrict__ src, int n_src, int spread, const RtF& Tf, float thr2, const Slab& slab,
                                                 const int* hint, int* corr, float* __restrict__ sqd, int* __restrict__ need_state, int* __restrict__ need_list,
                                                 int* __restrict__ need_count, int gt) {
  const int i = gt / spread;
  const bool grouped = group_search(spread);
  const bool leader = (gt & (spread - 1)) == 0;
  if (i >= n_src || (!grouped && !leader)) return;
  const float4 p = __ldg(&src[i]);
  float qx, qy, qz;
  transform_f(Tf.m, p.x, p.y, p.z, qx, qy, qz);
  Best1 top;
  top.reset(1, thr2);
  if (g_tile_dbg) {  // debug statistics (rgc_debug_correspond_stats): cycles, nodes, lookups, candidates
    SearchStats st{0, 0, 0};
    const long long t0 = clock64();
    if (slab_owns(slab, qx, qy, qz)) knn_search(tgt, qx, qy, qz, 1, thr2, hint ? hint[i] : -1, top, &st);
    long long* o = g_tile_dbg + (size_t)i * 4;
    o[0] = clock64() - t0;
    o[1] = st.nodes;
    o[2] = st.lookups;
    o[3] = st.candidates;
  } else if (slab_owns(slab, qx, qy, qz)) {
    nn1_spread(tgt, qx, qy, qz, thr2, hint ? hint[i] : -1, grouped ? spread : 1, top);
  }
  if (!leader) return;
  const int pos = (top.id0 >= 0 && top.d0 < thr2) ? top.id0 : -1;
  corr[i] = pos;
  sqd[i] = top.d0;
  if (need_state) {
    const bool claim = pos >= 0 && need_state[pos] == 0 && atomicCAS(&need_state[pos], 0, 1) == 0;
    // one atomicAdd per group of lanes that reach this point together
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, claim);
    if (m) {
      const int lane = threadIdx.x & 31, leader_lane = __ffs(m) - 1;
      int base = 0;
      if (lane == leader_lane) base = atomicAdd(need_count, __popc(m));
      base = __shfl_sync(act, base, leader_lane);
      if (claim) need_list[base + __popc(m & ((1u << lane) - 1u))] = pos;
    }
  }
}
__global__ void __launch_bounds__(kThreads, RGC_CORR_MINB) k_correspond(GridView tgt, const float4* __restrict__ src, int n_src, int spread, RtF Tf, float thr2, Slab slab,
                                                            const int* hint, int* corr, float* __restrict__ sqd, int* __restrict__ need_state,
                                                            int* __restrict__ need_list, int* __restrict__ need_count) {
  pdl_enter();
  correspond_query(tgt, src, n_src, spread, Tf, thr2, slab, hint, corr, sqd, need_state, need_list, need_count, blockIdx.x * blockDim.x + threadIdx.x);
}

// Mahalanobis part of update_correspondences + linearize (fast_gicp_impl.hpp:139-211), fused:
// per correspondence M = (C_B + R C_A R^T)^-1 (stored for compute_error), e^T M e, the 21 unique
// entries of H = J^T M J and b = J^T M e, reduced deterministically.  Streams p, C_A, corr; gathers
// q, C_B; writes M: 184 algorithmic bytes per source point -> HBM-bound once the batch exceeds L2.
//
// The loop over a thread's points (local index il = first, first + stride, ... < n; global = base + il) is
// software-pipelined: corr[] is read two points ahead and p, C_A and the dependent gathers q, C_B one point
// ahead, so the loads of the next point are in flight while the ~400 fp64 instructions of the current one
// issue (the unpipelined loop ran the memory phase and the arithmetic phase of all 16 warps of an SM back to
// back: 52 % of HBM peak at 2 M points, profiles/README.md).  The terms are added in ascending il, as before.
#ifndef RGC_LIN_MINB
#define RGC_LIN_MINB 4
#endif
struct LinPoint {
  float4 p, q;
  Sym3 CA, CB;
};
__device__ __forceinline__ void lin_load(LinPoint& d, const float4* __restrict__ tgt_pts, const float4* __restrict__ src, const double* __restrict__ src_cov,
                                         const double* __restrict__ tgt_cov, int i, int pos) {
  d.p = __ldg(&src[i]);
  d.q = __ldg(&tgt_pts[pos]);
  d.CA = load_sym3(src_cov, i);
  d.CB = load_sym3(tgt_cov, pos);
}
// staged operands of one correspondence: 8 x 16 bytes per thread and stage, chunk-major so that the 128-bit
// shared-memory reads of a warp are conflict-free: [stage][chunk][thread]
//   chunk 0: p   1: q   2-4: C_A   5-7: C_B
constexpr int kLinStages = 2;
__device__ __forceinline__ void lin_stage_issue(float4 (*st)[kThreads], const float4* __restrict__ tgt_pts, const float4* __restrict__ src,
                                                const double* __restrict__ src_cov, const double* __restrict__ tgt_cov, int i, int pos) {
  const int t = threadIdx.x;
  cp_async16(&st[0][t], &src[i], false);
  cp_async16(&st[1][t], &tgt_pts[pos], false);
  const float4* ca = reinterpret_cast<const float4*>(src_cov + (size_t)i * 6);
  const float4* cb = reinterpret_cast<const float4*>(tgt_cov + (size_t)pos * 6);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    cp_async16(&st[2 + c][t], ca + c, false);
    cp_async16(&st[5 + c][t], cb + c, false);
  }
}
__device__ __forceinline__ Sym3 lin_stage_sym3(float4 (*st)[kThreads], int c0) {
  const int t = threadIdx.x;
  const double2 a = *reinterpret_cast<const double2*>(&st[c0][t]), b = *reinterpret_cast<const double2*>(&st[c0 + 1][t]),
                c = *reinterpret_cast<const double2*>(&st[c0 + 2][t]);
  return Sym3{a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ void linearize_points(const float4* __restrict__ tgt_pts, const float4* __restrict__ src, const double* __restrict__ src_cov,
                                                 const double* __restrict__ tgt_cov, const int* __restrict__ corr, double* __restrict__ maha, const Rt& Td,
                                                 int want_hb, int base, int first, int stride, int n, double* acc) {
#if RGC_ASYNC_STAGE
  // the next point's 128 bytes are copied global -> shared asynchronously (no registers in flight) while the
  // ~400 fp64 instructions of the current one issue; corr[] is read two points ahead as before.  The terms
  // are added in ascending il, as before: same bits.
  __shared__ __align__(16) float4 stage[kLinStages][8][kThreads];
  int il = first;
  int c1 = il < n ? __ldg(&corr[base + il]) : -1;
  int c2 = il + stride < n ? __ldg(&corr[base + il + stride]) : -1;
  int cur = 0;
  if (c1 >= 0) lin_stage_issue(stage[0], tgt_pts, src, src_cov, tgt_cov, base + il, c1);
  cp_async_commit();
  while (il < n) {
    const int iln = il + stride;
    const int c3 = (iln + stride < n) ? __ldg(&corr[base + iln + stride]) : -1;
    if (c2 >= 0) lin_stage_issue(stage[cur ^ 1], tgt_pts, src, src_cov, tgt_cov, base + iln, c2);
    cp_async_commit();
    cp_async_wait<1>();  // everything but the group just committed has landed: the current point
    if (c1 >= 0) {
      float4(*st)[kThreads] = stage[cur];
      const float4 p = st[0][threadIdx.x], q = st[1][threadIdx.x];
      const Sym3 M = gicp_mahalanobis(Td, lin_stage_sym3(st, 2), lin_stage_sym3(st, 5));
      store_sym3(maha, base + il, M);
      if (want_hb)
        gicp_point_terms(Td, M, p.x, p.y, p.z, q.x, q.y, q.z, acc);
      else
        acc[0] = dadd(acc[0], gicp_error_term(Td, M, p.x, p.y, p.z, q.x, q.y, q.z));
      acc[kAccN] += 1.0;
    }
    cur ^= 1;
    c1 = c2;
    c2 = c3;
    il = iln;
  }
  cp_async_wait<0>();
#else
  int il = first;
  int c1 = il < n ? __ldg(&corr[base + il]) : -1;
  int c2 = il + stride < n ? __ldg(&corr[base + il + stride]) : -1;
  LinPoint cur, nxt;
  if (c1 >= 0) lin_load(cur, tgt_pts, src, src_cov, tgt_cov, base + il, c1);
  while (il < n) {
    const int iln = il + stride;
    const int c3 = (iln + stride < n) ? __ldg(&corr[base + iln + stride]) : -1;
    if (c2 >= 0) lin_load(nxt, tgt_pts, src, src_cov, tgt_cov, base + iln, c2);
    if (c1 >= 0) {
      const Sym3 M = gicp_mahalanobis(Td, cur.CA, cur.CB);
      store_sym3(maha, base + il, M);
      if (want_hb)
        gicp_point_terms(Td, M, cur.p.x, cur.p.y, cur.p.z, cur.q.x, cur.q.y, cur.q.z, acc);
      else
        acc[0] = dadd(acc[0], gicp_error_term(Td, M, cur.p.x, cur.p.y, cur.p.z, cur.q.x, cur.q.y, cur.q.z));
      acc[kAccN] += 1.0;
    }
    cur = nxt;
    c1 = c2;
    c2 = c3;
    il = iln;
  }
#endif
}
__global__ void __launch_bounds__(kThreads, RGC_LIN_MINB) k_linearize(const float4* __restrict__ tgt_pts, const float4* __restrict__ src, const double* __restrict__ src_cov,
                                                           const double* __restrict__ tgt_cov, int n_src, Rt Td, int want_hb, const int* __restrict__ corr,
                                                           double* __restrict__ maha, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                           double* __restrict__ result, DoneFlag done, int* __restrict__ zero_me) {
  pdl_enter();
  double acc[kLinN];
#pragma unroll
  for (int j = 0; j < kLinN; j++) acc[j] = 0.0;
  linearize_points(tgt_pts, src, src_cov, tgt_cov, corr, maha, Td, want_hb, 0, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, n_src, acc);
  grid_reduce<kLinN>(acc, partials, ticket, result, done, zero_me);
}

// fast_gicp_impl.hpp:214-237 — correspondences and M frozen from the last k_linearize.  Same pipelining.
struct CePoint {
  float4 p, q;
  Sym3 M;
};
__device__ __forceinline__ void compute_error_points(const float4* __restrict__ tgt_pts, const float4* __restrict__ src, const int* __restrict__ corr,
                                                     const double* __restrict__ maha, const Rt& Td, int base, int first, int stride, int n, double* acc) {
#if RGC_ASYNC_STAGE
  // same staging as linearize_points: p, q and the frozen M of the next point (80 bytes) are copied global -> shared
  // asynchronously while the current one is evaluated
  __shared__ __align__(16) float4 stage[2][5][kThreads];
  const int t = threadIdx.x;
  auto issue = [&](float4 (*st)[kThreads], int i, int pos) {
    cp_async16(&st[0][t], &src[i], false);
    cp_async16(&st[1][t], &tgt_pts[pos], false);
    const float4* m = reinterpret_cast<const float4*>(maha + (size_t)i * 6);
#pragma unroll
    for (int c = 0; c < 3; c++) cp_async16(&st[2 + c][t], m + c, false);
  };
  int il = first;
  int c1 = il < n ? __ldg(&corr[base + il]) : -1;
  int c2 = il + stride < n ? __ldg(&corr[base + il + stride]) : -1;
  int cur = 0;
  if (c1 >= 0) issue(stage[0], base + il, c1);
  cp_async_commit();
  while (il < n) {
    const int iln = il + stride;
    const int c3 = (iln + stride < n) ? __ldg(&corr[base + iln + stride]) : -1;
    if (c2 >= 0) issue(stage[cur ^ 1], base + iln, c2);
    cp_async_commit();
    cp_async_wait<1>();
    if (c1 >= 0) {
      float4(*st)[kThreads] = stage[cur];
      const float4 p = st[0][t], q = st[1][t];
      acc[0] = dadd(acc[0], gicp_error_term(Td, lin_stage_sym3(st, 2), p.x, p.y, p.z, q.x, q.y, q.z));
    }
    cur ^= 1;
    c1 = c2;
    c2 = c3;
    il = iln;
  }
  cp_async_wait<0>();
#else
  int il = first;
  int c1 = il < n ? __ldg(&corr[base + il]) : -1;
  int c2 = il + stride < n ? __ldg(&corr[base + il + stride]) : -1;
  CePoint cur, nxt;
  auto load = [&](CePoint& d, int i, int pos) {
    d.p = __ldg(&src[i]);
    d.q = __ldg(&tgt_pts[pos]);
    d.M = load_sym3(maha, i);
  };
  if (c1 >= 0) load(cur, base + il, c1);
  while (il < n) {
    const int iln = il + stride;
    const int c3 = (iln + stride < n) ? __ldg(&corr[base + iln + stride]) : -1;
    if (c2 >= 0) load(nxt, base + iln, c2);
    if (c1 >= 0) acc[0] = dadd(acc[0], gicp_error_term(Td, cur.M, cur.p.x, cur.p.y, cur.p.z, cur.q.x, cur.q.y, cur.q.z));
    cur = nxt;
    c1 = c2;
    c2 = c3;
    il = iln;
  }
#endif
}
__global__ void __launch_bounds__(kThreads) k_compute_error(const float4* __restrict__ tgt_pts, const float4* __restrict__ src, int n_src, Rt Td,
                                                            const int* __restrict__ corr, const double* __restrict__ maha,
                                                            double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ result,
                                                            DoneFlag done) {
  pdl_enter();
  double acc[1] = {0.0};
  compute_error_points(tgt_pts, src, corr, maha, Td, 0, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, n_src, acc);
  grid_reduce<1>(acc, partials, ticket, result, done);
}

// One launch for the two independent jobs of an LM trial step at pose xi (lsq_registration_impl.hpp:141-143 and,
// if the step is accepted, :127 of the next iteration): blocks [0, ce_blocks) evaluate compute_error(xi) on the
// frozen correspondences / Mahalanobis matrices of the last linearize — the same virtual grid, thread mapping and
// reduction order as k_compute_error, hence the same bits — while the remaining blocks search the correspondences
// of linearize(xi) (hinted by the frozen ones) into the second buffer set.  Run back to back the error kernel was
// a 10 us latency chain in front of a 55 us one; side by side it disappears behind the search.
__global__ void __launch_bounds__(kThreads, RGC_CORR_MINB) k_trial_step(GridView tgt, const float4* __restrict__ src, int n_src, int spread, RtF Tf, float thr2, Slab slab,
                                                            const int* hint, int* corr, float* __restrict__ sqd, int* __restrict__ need_state,
                                                            int* __restrict__ need_list, int* __restrict__ need_count, int ce_blocks, Rt Td,
                                                            const double* __restrict__ ce_maha, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                            double* __restrict__ ce_result, DoneFlag ce_done) {
  pdl_enter();
  if ((int)blockIdx.x < ce_blocks) {
    double acc[1] = {0.0};
    compute_error_points(reinterpret_cast<const float4*>(tgt.pts), src, hint, ce_maha, Td, 0, blockIdx.x * blockDim.x + threadIdx.x, ce_blocks * blockDim.x, n_src, acc);
    grid_reduce_at<1>(acc, partials, blockIdx.x, (unsigned)ce_blocks, ticket, ce_result, ce_done);
    return;
  }
  correspond_query(tgt, src, n_src, spread, Tf, thr2, slab, hint, corr, sqd, need_state, need_list, need_count,
                   ((int)blockIdx.x - ce_blocks) * blockDim.x + threadIdx.x);
}

// pcl::Registration::getFitnessScore: [sum d2, count] over 1-NN d2 <= max_range.  `hint` (nullable): the
// correspondences of the last linearize — the final pose is next to that one, so the old neighbour is a tight
// first bound (any real point is a valid bound: the result is the same exact nearest neighbour)
__global__ void __launch_bounds__(kThreads, RGC_CORR_MINB) k_fitness(GridView tgt, const float4* __restrict__ src, int n_src, int spread, RtF Tf, double max_range, Slab slab,
                                                      const int* __restrict__ hint, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                      double* __restrict__ result, DoneFlag done) {
  double acc[2] = {0.0, 0.0};
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gt / spread;
  const bool grouped = group_search(spread);
  const bool leader = (gt & (spread - 1)) == 0;
  if (i < n_src && (grouped || leader)) {
    const float4 p = __ldg(&src[i]);
    float qx, qy, qz;
    transform_f(Tf.m, p.x, p.y, p.z, qx, qy, qz);
    Best1 top;
    top.reset(1, INFINITY);
    if (slab_owns(slab, qx, qy, qz)) nn1_spread(tgt, qx, qy, qz, INFINITY, hint ? __ldg(&hint[i]) : -1, grouped ? spread : 1, top);
    if (leader && top.id0 >= 0 && (double)top.d0 <= max_range) {
      acc[0] = (double)top.d0;
      acc[1] = 1.0;
    }
  }
  grid_reduce<2>(acc, partials, ticket, result, done);
}

__global__ void __launch_bounds__(256) k_transform_out(const float4* __restrict__ src_sorted, int n, RtF Tf, float4* __restrict__ out_orig_order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = src_sorted[i];
  float x, y, z;
  transform_f(Tf.m, p.x, p.y, p.z, x, y, z);
  out_orig_order[__float_as_int(p.w)] = make_float4(x, y, z, 1.0f);
}

// correspondences in the caller's index space
__global__ void __launch_bounds__(256) k_corr_to_orig(const float4* __restrict__ src_sorted, const float4* __restrict__ tgt_sorted, const int* __restrict__ corr,
                                                      const float* __restrict__ sqd, int n, int* __restrict__ corr_out, float* __restrict__ sqd_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int o = __float_as_int(src_sorted[i].w);
  int c = corr[i];
  corr_out[o] = c >= 0 ? __float_as_int(tgt_sorted[c].w) : -1;
  sqd_out[o] = sqd[i];
}

// k-major neighbour positions (sorted order) -> row-major original indices in the caller's order
__global__ void __launch_bounds__(256) k_nbr_to_orig(const float4* __restrict__ sorted, const int* __restrict__ nbr, int n, int k, int* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int o = __float_as_int(sorted[t].w);
  for (int j = 0; j < k; j++) {
    const int p = nbr[(size_t)j * n + t];
    out[(size_t)o * k + j] = p >= 0 ? __float_as_int(sorted[p].w) : -1;
  }
}

// covariances between the caller's layout (4x4 doubles, original order) and ours (6 doubles, sorted)
__global__ void __launch_bounds__(256) k_cov_import(const float4* __restrict__ sorted, int n, const double* __restrict__ m4x4, double* __restrict__ cov6) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* m = m4x4 + (size_t)__float_as_int(sorted[i].w) * 16;
  // column-major or row-major is irrelevant for the symmetric 3x3 block; take the upper triangle
  double* o = cov6 + (size_t)i * 6;
  o[0] = m[0]; o[1] = m[4]; o[2] = m[8]; o[3] = m[5]; o[4] = m[9]; o[5] = m[10];
}
__global__ void __launch_bounds__(256) k_cov_export(const float4* __restrict__ sorted, int n, const double* __restrict__ cov6, double* __restrict__ m4x4) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* m = m4x4 + (size_t)__float_as_int(sorted[i].w) * 16;
  const double* c = cov6 + (size_t)i * 6;
  m[0] = c[0]; m[1] = c[1]; m[2] = c[2]; m[3] = 0.0;
  m[4] = c[1]; m[5] = c[3]; m[6] = c[4]; m[7] = 0.0;
  m[8] = c[2]; m[9] = c[4]; m[10] = c[5]; m[11] = 0.0;
  m[12] = 0.0; m[13] = 0.0; m[14] = 0.0; m[15] = 0.0;
}

// summed partial results (device buffer, after the all-reduce of a sharded registration) -> the mapped host
// result area: src[0, n_a) -> dst_a, src[n_a, n_a + n_b) -> dst_b, then the completion word
__global__ void __launch_bounds__(64) k_publish(const double* __restrict__ src, int n_a, double* __restrict__ dst_a, int n_b, double* __restrict__ dst_b, DoneFlag done) {
  const int t = threadIdx.x;
  if (t < n_a) dst_a[t] = src[t];
  if (t < n_b) dst_b[t] = src[n_a + t];
  __threadfence_system();
  __syncthreads();
  if (t == 0 && done.flag) *reinterpret_cast<volatile unsigned long long*>(done.flag) = done.value;
  __threadfence_system();
}

// ---- all-reduce + publish over NVLink peer memory (config C5), one launch per LM step.
// Every rank owns a mailbox (device memory, mapped into the other ranks' address spaces through CUDA IPC):
//   mail[parity][sender][2 * 64] 64-bit words, each = (sequence number << 32) | one 32-bit half of a double.
// Flag-in-data, as in NCCL's low-latency protocol: an aligned 8-byte store is indivisible, so a word whose upper half
// shows this call's sequence number carries this call's data — no fence between data and flag, one NVLink trip.
// The block (a) stores the two halves of each of this rank's partial sums into slot [parity][rank] of EVERY rank's
// mailbox (plain stores over NVLink), (b) polls its OWN mailbox until every sender's words show the sequence number,
// (c) adds the world's partials in rank order — the same order on every rank, so all ranks get bit-identical totals
// and take identical LM steps — and (d) writes the totals back in place and, when asked, into the mapped host result
// area followed by the completion word the host polls.  (A first version — data, __threadfence_system(), flag
// word, fence — took 13.1 us per call on two GPUs against ncclAllReduce's 9.6 us.)
// Two parities: a rank can be one call ahead of a peer (it starts call s + 1 once it has everybody's call-s
// data, while a peer may still be reading its own copy of call s), never two.  A peer that never arrives turns
// the totals into NaN after ~2 s instead of hanging the stream.
constexpr int kPeerMax = 16;
constexpr int kPeerSlot = 64;
struct PeerSet {
  unsigned long long* mail[kPeerMax];
};
__global__ void __launch_bounds__(64) k_peer_allreduce(double* __restrict__ buf, int n, PeerSet ps, int rank, int world, unsigned long long seq, int n_a,
                                                       double* __restrict__ dst_a, int n_b, double* __restrict__ dst_b, DoneFlag done) {
  pdl_enter();
  const int t = threadIdx.x;
  const size_t par_off = (size_t)(seq & 1ull) * kPeerMax * 2 * kPeerSlot;
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  if (t < n) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(buf[t]);
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    const size_t at = par_off + (size_t)rank * 2 * kPeerSlot + 2 * t;
    for (int r = 0; r < world; r++) {
      volatile unsigned long long* m = ps.mail[r] + at;
      m[0] = w0;
      m[1] = w1;
    }
    const volatile unsigned long long* mine = ps.mail[rank] + par_off + 2 * t;
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    double sum = 0.0;
    bool ok = true;
    for (int r = 0; r < world; r++) {
      unsigned long long a, b;
      for (;;) {
        a = mine[(size_t)r * 2 * kPeerSlot];
        b = mine[(size_t)r * 2 * kPeerSlot + 1];
        if ((a & 0xffffffff00000000ull) == tag && (b & 0xffffffff00000000ull) == tag) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 2000000000ull) {
          ok = false;
          break;
        }
      }
      sum = dadd(sum, __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32))));
    }
    if (!ok) sum = __longlong_as_double(0x7ff8000000000000ll);
    buf[t] = sum;
    if (dst_a && t < n_a) dst_a[t] = sum;
    if (dst_b && t >= n_a && t < n_a + n_b) dst_b[t - n_a] = sum;
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0 && done.flag) *reinterpret_cast<volatile unsigned long long*>(done.flag) = done.value;
  __threadfence_system();
}

// config C5: the points of a raw cloud whose `axis` coordinate lies in [lo, hi) (a rank's slab + halo), kept in
// input order: per-block counts -> k_vg_scan_blocks -> ordered scatter (+ the index of each kept point)
constexpr int kSlabItems = 8;
__device__ __forceinline__ bool slab_keep(const unsigned char* __restrict__ raw, size_t stride, int i, int n, int axis, float lo, float hi) {
  if (i >= n) return false;
  const float v = reinterpret_cast<const float*>(raw + (size_t)i * stride)[axis];
  return v >= lo && v < hi;
}
__global__ void __launch_bounds__(256) k_slab_count(const unsigned char* __restrict__ raw, size_t stride, int n, int axis, float lo, float hi,
                                                    unsigned int* __restrict__ block_counts) {
  const int base = blockIdx.x * 256 * kSlabItems;
  int c = 0;
#pragma unroll
  for (int u = 0; u < kSlabItems; u++) c += slab_keep(raw, stride, base + u * 256 + threadIdx.x, n, axis, lo, hi) ? 1 : 0;
  __shared__ int ws[8];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; w++) t += ws[w];
    block_counts[blockIdx.x] = (unsigned)t;
  }
}
__global__ void __launch_bounds__(256) k_slab_scatter(const unsigned char* __restrict__ raw, size_t stride, int n, int axis, float lo, float hi,
                                                      const unsigned int* __restrict__ block_offsets, float4* __restrict__ out, int* __restrict__ index_out) {
  const int base = blockIdx.x * 256 * kSlabItems;
  __shared__ int warp_cnt[8];
  __shared__ int pass_base;
  if (threadIdx.x == 0) pass_base = (int)block_offsets[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = 0; u < kSlabItems; u++) {
    const int i = base + u * 256 + threadIdx.x;
    const bool keep = slab_keep(raw, stride, i, n, axis, lo, hi);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = pass_base;
    for (int w = 0; w < warp; w++) before += warp_cnt[w];
    if (keep) {
      const int pos = before + __popc(bal & ((1u << lane) - 1u));
      const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
      out[pos] = make_float4(p[0], p[1], p[2], 1.0f);
      if (index_out) index_out[pos] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; w++) t += warp_cnt[w];
      pass_base += t;
    }
    __syncthreads();
  }
}

// per-point int flags, sorted order -> the caller's order (test hook for the on-demand covariance state)
__global__ void __launch_bounds__(256) k_flags_to_orig(const float4* __restrict__ sorted, int n, const int* __restrict__ flags, int fill, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[__float_as_int(sorted[i].w)] = flags ? flags[i] : fill;
}

}  // namespace rgc
