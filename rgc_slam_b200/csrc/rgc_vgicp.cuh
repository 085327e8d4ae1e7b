// rgc_vgicp.cuh — voxelised GICP (fast_gicp::FastVGICP), the variant RGC_odometer.cpp:998 instantiates
// (SURVEY.md §8f N1).  Included by rgc_gicp.cu only.
//
//   GaussianVoxelMap::create_voxelmap   rgc_slam/include/fast_gicp/gicp/fast_vgicp_voxel.hpp:129-156
//   voxel_coord = floor(x / res - 0.5)  :158-160
//   neighbor_offsets DIRECT1/7/27       :10-44
//   update_correspondences / linearize / compute_error   impl/fast_vgicp_impl.hpp:73-204
//
// Voxel map build: target points are keyed by their voxel, (key, original index) pairs are sorted
// with the stable LSD radix sort of the cloud build, and one WARP per voxel adds the means and
// covariances of its run (lane-strided, ascending original index, fixed shuffle tree): deterministic,
// no atomics on the sums.  Voxels live in an open-addressing table keyed by the compact voxel key; a
// query is one probe per offset.
#pragma once
#include "rgc_kernels.cuh"

namespace rgc {

struct VoxGeom {
  double res;        // voxel_resolution_
  int lo[3];         // smallest voxel coordinate of the target per axis
  int bits[3];       // bits per axis of the compact key
  int dim[3];        // hi - lo + 1
};

struct VoxelSlot {  // 88 bytes
  unsigned long long key;
  int num;
  int pad;
  double mean[3];
  double cov[6];
};

struct VoxelMapView {
  const VoxelSlot* slots;
  uint32_t mask, shift;
  VoxGeom geom;
};

__device__ __forceinline__ int voxel_coord_d(double x, double res) { return (int)floor(x / res - 0.5); }

// compact key of voxel (cx,cy,cz), or ~0 when outside the target's voxel bounding box
__device__ __forceinline__ unsigned long long voxel_key(const VoxGeom& g, int cx, int cy, int cz) {
  const int x = cx - g.lo[0], y = cy - g.lo[1], z = cz - g.lo[2];
  if ((unsigned)x >= (unsigned)g.dim[0] || (unsigned)y >= (unsigned)g.dim[1] || (unsigned)z >= (unsigned)g.dim[2]) return ~0ull;
  return ((unsigned long long)z << (g.bits[0] + g.bits[1])) | ((unsigned long long)y << g.bits[0]) | (unsigned long long)x;
}

__device__ __forceinline__ int voxel_find(const VoxelMapView& v, unsigned long long key) {
  if (key == ~0ull) return -1;
  uint32_t h = slot_of(key, v.shift);
  for (;;) {
    const unsigned long long k = __ldg(&v.slots[h].key);
    if (k == key) return (int)h;
    if (k == ~0ull) return -1;
    h = (h + 1) & v.mask;
  }
}

// keys / values in ORIGINAL index order, so the stable sort keeps ascending original index per voxel
__global__ void __launch_bounds__(256) k_vox_keys(const float4* __restrict__ sorted, int n, VoxGeom g, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float4 p = sorted[t];
  const int o = __float_as_int(p.w);
  keys[o] = voxel_key(g, voxel_coord_d((double)p.x, g.res), voxel_coord_d((double)p.y, g.res), voxel_coord_d((double)p.z, g.res));
  vals[o] = (uint32_t)o;
}

// heads of the runs of equal keys: count them and append their positions (unordered) to `heads`
__global__ void __launch_bounds__(256) k_vox_count(const uint64_t* __restrict__ keys, int n, unsigned int* __restrict__ counter, int* __restrict__ heads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool head = i < n && (i == 0 || keys[i] != keys[i - 1]);
  const unsigned bal = __ballot_sync(0xffffffffu, head);
  if (!bal) return;
  const int lane = threadIdx.x & 31;
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(counter, (unsigned)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (head) heads[base + __popc(bal & ((1u << lane) - 1u))] = i;
}

// one WARP per voxel (fast_vgicp_voxel.hpp:105-122 additive, :80-101 multiplicative): lane l adds the
// run's points l, l+32, ... in ascending original index, then the 32 partial sums are combined by a
// fixed shuffle tree — deterministic, and the long runs of the dense near-sensor voxels (thousands of
// points) no longer serialise on one thread (the first version: 2.6 ms for the C2 target).
__global__ void __launch_bounds__(128) k_vox_reduce(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n, const int* __restrict__ heads,
                                                    int nv, const float4* __restrict__ sorted, const int* __restrict__ inv, const double* __restrict__ cov6,
                                                    int mode, VoxelSlot* __restrict__ slots, uint32_t mask, uint32_t shift) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nv) return;
  const int start = heads[w];
  const uint64_t key = keys[start];
  // run end: first index > start whose key differs (keys are sorted): binary search, uniform per warp
  int lo = start + 1, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] == key) lo = mid + 1; else hi = mid;
  }
  const int end = lo;
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // mean xyz, cov xx xy xz yy yz zz
  for (int t = start + lane; t < end; t += 32) {
    const int pos = inv[vals[t]];
    const float4 p = sorted[pos];
    const Sym3 ci = load_sym3(cov6, pos);
    if (mode == 2) {  // MULTIPLICATIVE
      const Sym3 q = inv_sym3(ci);
      acc[3] += q.xx; acc[4] += q.xy; acc[5] += q.xz; acc[6] += q.yy; acc[7] += q.yz; acc[8] += q.zz;
      acc[0] += q.xx * (double)p.x + q.xy * (double)p.y + q.xz * (double)p.z;
      acc[1] += q.xy * (double)p.x + q.yy * (double)p.y + q.yz * (double)p.z;
      acc[2] += q.xz * (double)p.x + q.yz * (double)p.y + q.zz * (double)p.z;
    } else {  // ADDITIVE / ADDITIVE_WEIGHTED
      acc[0] += (double)p.x; acc[1] += (double)p.y; acc[2] += (double)p.z;
      acc[3] += ci.xx; acc[4] += ci.xy; acc[5] += ci.xz; acc[6] += ci.yy; acc[7] += ci.yz; acc[8] += ci.zz;
    }
  }
#pragma unroll
  for (int j = 0; j < 9; j++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o);
  if (lane != 0) return;
  const int num = end - start;
  double m[3] = {acc[0], acc[1], acc[2]};
  Sym3 c = {acc[3], acc[4], acc[5], acc[6], acc[7], acc[8]};
  if (mode == 2) {
    const Sym3 f = inv_sym3(c);
    const double m0 = f.xx * m[0] + f.xy * m[1] + f.xz * m[2], m1 = f.xy * m[0] + f.yy * m[1] + f.yz * m[2], m2 = f.xz * m[0] + f.yz * m[1] + f.zz * m[2];
    m[0] = m0; m[1] = m1; m[2] = m2;
    c = f;
  } else {
    const double dn = (double)num;
    m[0] /= dn; m[1] /= dn; m[2] /= dn;
    c.xx /= dn; c.xy /= dn; c.xz /= dn; c.yy /= dn; c.yz /= dn; c.zz /= dn;
  }
  uint32_t h = slot_of(key, shift);
  for (;;) {
    const unsigned long long prev = atomicCAS(&slots[h].key, ~0ull, (unsigned long long)key);
    if (prev == ~0ull) break;
    h = (h + 1) & mask;
  }
  VoxelSlot* s = &slots[h];
  s->num = num;
  s->mean[0] = m[0]; s->mean[1] = m[1]; s->mean[2] = m[2];
  s->cov[0] = c.xx; s->cov[1] = c.xy; s->cov[2] = c.xz; s->cov[3] = c.yy; s->cov[4] = c.yz; s->cov[5] = c.zz;
}

__device__ __forceinline__ void vgicp_offset(int method, int o, int& dx, int& dy, int& dz) {
  if (method == 2) {  // DIRECT1
    dx = dy = dz = 0;
  } else if (method == 1) {  // DIRECT7: (0,0,0) (1,0,0) (-1,0,0) (0,1,0) (0,-1,0) (0,0,1) (0,0,-1)
    dx = o == 1 ? 1 : (o == 2 ? -1 : 0);
    dy = o == 3 ? 1 : (o == 4 ? -1 : 0);
    dz = o == 5 ? 1 : (o == 6 ? -1 : 0);
  } else {  // DIRECT27: for i, j, k in 0..2: (i-1, j-1, k-1)
    dx = o / 9 - 1;
    dy = (o / 3) % 3 - 1;
    dz = o % 3 - 1;
  }
}

// acc += w * terms of one (point, voxel) correspondence; q is the voxel mean (double)
__device__ __forceinline__ void vgicp_terms(const Rt& T, const Sym3& M, float px, float py, float pz, const double* q, double w, int want_hb, double* acc) {
  double a[3];
  transform_d(T, (double)px, (double)py, (double)pz, a[0], a[1], a[2]);
  const double e[3] = {q[0] - a[0], q[1] - a[1], q[2] - a[2]};
  const double Mm[3][3] = {{M.xx, M.xy, M.xz}, {M.xy, M.yy, M.yz}, {M.xz, M.yz, M.zz}};
  double Me[3];
#pragma unroll
  for (int i = 0; i < 3; i++) Me[i] = Mm[i][0] * e[0] + Mm[i][1] * e[1] + Mm[i][2] * e[2];
  acc[0] += w * (e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2]);
  if (!want_hb) return;
  const double J[3][6] = {{0.0, -a[2], a[1], -1.0, 0.0, 0.0}, {a[2], 0.0, -a[0], 0.0, -1.0, 0.0}, {-a[1], a[0], 0.0, 0.0, 0.0, -1.0}};
  double MJ[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int c = 0; c < 6; c++) MJ[i][c] = Mm[i][0] * J[0][c] + Mm[i][1] * J[1][c] + Mm[i][2] * J[2][c];
  int o = 1;
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = r; c < 6; c++) {
      acc[o] += w * (J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c] + J[2][r] * MJ[2][c]);
      o++;
    }
#pragma unroll
  for (int r = 0; r < 6; r++) acc[22 + r] += w * (J[0][r] * Me[0] + J[1][r] * Me[1] + J[2][r] * Me[2]);
}

// update_correspondences + linearize (fast_vgicp_impl.hpp:73-180), one (source point, offset) pair per
// loop iteration, fixed order => deterministic sums
__global__ void __launch_bounds__(kThreads, 4) k_vgicp_linearize(VoxelMapView vm, const float4* __restrict__ src, const double* __restrict__ src_cov, int n_src,
                                                                 int method, int n_off, Rt Td, int want_hb, int* __restrict__ corr_slot,
                                                                 double* __restrict__ maha, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                                                 double* __restrict__ result, DoneFlag done) {
  double acc[kLinN];
#pragma unroll
  for (int j = 0; j < kLinN; j++) acc[j] = 0.0;
  const long long total = (long long)n_src * n_off;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n_off), o = (int)(idx % n_off);
    const float4 p = __ldg(&src[i]);
    double a[3];
    transform_d(Td, (double)p.x, (double)p.y, (double)p.z, a[0], a[1], a[2]);
    int dx, dy, dz;
    vgicp_offset(method, o, dx, dy, dz);
    const int slot = voxel_find(vm, voxel_key(vm.geom, voxel_coord_d(a[0], vm.geom.res) + dx, voxel_coord_d(a[1], vm.geom.res) + dy,
                                              voxel_coord_d(a[2], vm.geom.res) + dz));
    corr_slot[idx] = slot;
    if (slot < 0) continue;
    const VoxelSlot* vs = vm.slots + slot;
    const Sym3 CA = load_sym3(src_cov, i);
    const Sym3 CB = {vs->cov[0], vs->cov[1], vs->cov[2], vs->cov[3], vs->cov[4], vs->cov[5]};
    const Sym3 M = gicp_mahalanobis(Td, CA, CB);
    store_sym3(maha, (size_t)idx, M);
    const double q[3] = {vs->mean[0], vs->mean[1], vs->mean[2]};
    vgicp_terms(Td, M, p.x, p.y, p.z, q, sqrt((double)vs->num), want_hb, acc);
    acc[kAccN] += 1.0;
  }
  grid_reduce<kLinN>(acc, partials, ticket, result, done);
}

// fast_vgicp_impl.hpp:183-204: correspondences and Mahalanobis matrices frozen
__global__ void __launch_bounds__(kThreads) k_vgicp_compute_error(VoxelMapView vm, const float4* __restrict__ src, int n_src, int n_off, Rt Td,
                                                                  const int* __restrict__ corr_slot, const double* __restrict__ maha,
                                                                  double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ result,
                                                                  DoneFlag done) {
  double acc[1] = {0.0};
  const long long total = (long long)n_src * n_off;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int slot = __ldg(&corr_slot[idx]);
    if (slot < 0) continue;
    const int i = (int)(idx / n_off);
    const float4 p = __ldg(&src[i]);
    const VoxelSlot* vs = vm.slots + slot;
    const Sym3 M = load_sym3(maha, (size_t)idx);
    const double q[3] = {vs->mean[0], vs->mean[1], vs->mean[2]};
    vgicp_terms(Td, M, p.x, p.y, p.z, q, sqrt((double)vs->num), 0, acc);
  }
  grid_reduce<1>(acc, partials, ticket, result, done);
}

// test hook: dump the occupied voxels (unordered)
__global__ void __launch_bounds__(256) k_vox_dump(const VoxelSlot* __restrict__ slots, uint32_t nslots, VoxGeom g, unsigned int* __restrict__ counter,
                                                  int* __restrict__ coords, int* __restrict__ num, double* __restrict__ mean3, double* __restrict__ cov6,
                                                  unsigned int cap) {
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nslots || slots[h].key == ~0ull) return;
  const unsigned int o = atomicAdd(counter, 1u);
  if (o >= cap) return;
  const unsigned long long key = slots[h].key;
  coords[3 * o] = (int)(key & ((1ull << g.bits[0]) - 1)) + g.lo[0];
  coords[3 * o + 1] = (int)((key >> g.bits[0]) & ((1ull << g.bits[1]) - 1)) + g.lo[1];
  coords[3 * o + 2] = (int)(key >> (g.bits[0] + g.bits[1])) + g.lo[2];
  num[o] = slots[h].num;
  for (int d = 0; d < 3; d++) mean3[3 * o + d] = slots[h].mean[d];
  for (int d = 0; d < 6; d++) cov6[6 * o + d] = slots[h].cov[d];
}

}  // namespace rgc
