// rgc_gicp.cu — host side of librgc_gicp.so: context / memory pool, cloud build pipeline,
// the registration object and its LM driver, and the C-ABI declared in include/rgc_gicp.h.
//
// Reference behaviour mirrored here (paths under /root/reference/rgc_slam/include/fast_gicp/gicp/):
//   impl/fast_gicp_impl.hpp:72-112   setInputSource/Target caching, lazy covariances, align
//   impl/lsq_registration_impl.hpp:53-172  GN / LM outer loop, lambda schedule, convergence
// There is no CPU fallback anywhere in this file: every numeric step is a kernel launch.
#include <cuda_runtime.h>

#include <cfloat>
#include <climits>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "rgc_ctx.hpp"
#include "rgc_kernels.cuh"
#include "rgc_lm.hpp"
#include "../../include/rgc_mapping.h"
#include "../../include/rgc_preprocess.h"
#include "rgc_mapping.cuh"
#include "rgc_preprocess.cuh"
#include "rgc_vgicp.cuh"
#include "../../include/rgc_batch.h"
#include "rgc_batch.cuh"

using namespace rgc;

// Launch `kern` so that it may be scheduled while the kernel before it on `st` is still draining (programmatic
// dependent launch).  Only for kernels that start with pdl_enter() (rgc_common.cuh); with c->pdl off (RGC_NO_PDL=1)
// this is a plain launch.  Arguments are converted to the kernel's parameter types by cudaLaunchKernelEx.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(const rgc_ctx* c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = c->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

#include "rgc_comm.inl"

// RGC_TRACE=1: wall-clock trace of the host side of the build pipeline (debug aid)
static bool trace_on() {
  static int v = -1;
  if (v < 0) v = std::getenv("RGC_TRACE") ? 1 : 0;
  return v == 1;
}
struct Tracer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!trace_on()) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[rgc trace] %-28s %9.1f us\n", what, std::chrono::duration<double, std::micro>(t1 - t0).count());
    t0 = t1;
  }
};

// ------------------------------------------------------------------------------------------------
// A cloud build in flight.  The build has two points where the HOST needs a number from the device (the
// bounding box -> grid geometry and number of sort passes; the per-level cell counts -> table sizes).  With
// deferred builds (the default) a setter issues the work up to the first of them and returns; the rest is
// driven by the next call that needs the cloud (reg_drain), which interleaves the builds of source and
// target so that each host wait overlaps the other cloud's device work (they used to run back to back:
// 0.6 ms of prologue for a cold align instead of ~0.35).
struct BuildJob {
  Scratch tmp;
  int stage = 0;  // 1: ingest issued, waiting for the bounding box; 2: sort issued, waiting for the level counts
  bool spec = false;      // phase 2 was issued with the previous cloud's grid geometry, before the bounding box was known
  int max_bits = kMaxBits;
  int lane = 0;
  int n = 0, cloud_bits = 0, n_clouds = 0;
  uint64_t key = 0;
  float cell = 0.f;
  float4* orig = nullptr;
  uint64_t *keys_a = nullptr, *keys_b = nullptr, *kin = nullptr;
  uint32_t *vals_a = nullptr, *vals_b = nullptr, *vin = nullptr, *hist = nullptr, *d_counts = nullptr;
  float* h_slot = nullptr;     // pinned + mapped: 6 x kBboxBlocks floats, then kMaxLevels counters and the sort's error word
  float* d_slot = nullptr;     // device alias of h_slot: the kernels write those numbers straight into it
  cudaEvent_t ready = nullptr;  // the host copy the next phase waits for has landed
  // speculative level tables (sizes of the previous cloud of the lane): cleared on the lane's aux stream during the
  // sort, filled right behind it; the cell counts are taken on the aux stream meanwhile
  bool spec_tables = false;
  size_t spec_slots[kMaxLevels] = {};
  cudaEvent_t cleared = nullptr;
  explicit BuildJob(rgc_ctx* c) : tmp(c) {}
};

struct Cloud {
  std::unique_ptr<BuildJob> job;  // non-null while the build is in flight (valid is false until it completes)
  int n = 0;
  uint64_t key = 0;
  bool valid = false;
  float4* sorted = nullptr;
  int* inv = nullptr;  // original index -> sorted position
  GridSlot* tables = nullptr;
  GridView view{};
  double* cov = nullptr;  // 6 doubles per sorted point
  // on-demand mode (target of FastGICP): cov[] is valid only where cov_state[] == 1; has_cov stays false,
  // so every consumer that needs ALL covariances (getTargetCovariances, the voxel map, a swap) still
  // triggers the whole-cloud pass
  int* cov_state = nullptr;
  bool lazy_cov = false;
  bool has_cov = false;
  bool cov_speculative = false;  // computed at set_input time, before any align has used them
  int cov_k = 0, cov_method = 0;
  float build_ms = 0, knn_ms = 0, cov_ms = 0;
  float bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};
  int lane_built = 0;  // lane (stream) the build and the ahead-of-time covariances were issued on
  // multi-cloud grid (batched registration): n_clouds clouds concatenated; cloud c = sorted positions
  // [h_off[c], h_off[c + 1]); d_off is the device copy, tiles the tile list of the self-kNN kernel
  int n_clouds = 0;
  std::vector<int> h_off;
  int* d_off = nullptr;
  TileDesc* d_tiles = nullptr;
  int n_tiles = 0;
  // stage timing: build begin / end, kNN begin, kNN end (= covariance begin), covariance end.
  // Read back lazily (cloud_times) so that nothing here makes the host wait for the device.
  // ev[5]: the sorted points are final (recorded behind the radix sort's gather pass)
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool build_timed = false, cov_timed = false;
};

static void cloud_times(Cloud& cl) {
  if (!cl.valid) return;
  if (cl.build_timed && cudaEventSynchronize(cl.ev[1]) == cudaSuccess) {
    cudaEventElapsedTime(&cl.build_ms, cl.ev[0], cl.ev[1]);
    cl.build_timed = false;
  }
  if (cl.cov_timed && cudaEventSynchronize(cl.ev[4]) == cudaSuccess) {
    cudaEventElapsedTime(&cl.knn_ms, cl.ev[2], cl.ev[3]);
    cudaEventElapsedTime(&cl.cov_ms, cl.ev[3], cl.ev[4]);
    cl.cov_timed = false;
  }
}

// runs the enclosed work on lane 1 (rgc_ctx.hpp); the main stream picks it up with join_side()
struct SideLane {
  rgc_ctx* c;
  bool on;
  SideLane(rgc_ctx* c_, bool on_) : c(c_), on(on_ && c_->overlap) {
    if (on) c->switch_lane(1);
  }
  ~SideLane() {
    if (!on) return;
    cudaEventRecord(c->join_ev, c->stream);
    c->side_pending = true;
    c->switch_lane(0);
  }
};
static int join_side(rgc_ctx* c) {
  if (c->side_pending) {
    CK(c, cudaStreamWaitEvent(c->stream, c->join_ev, 0));
    c->side_pending = false;
  }
  return RGC_OK;
}

// The main stream may read the POINTS of a source cloud prepared on lane 1 as soon as they are sorted (cl.ev[5], or
// the end of the build, cl.ev[1], with RGC_JOIN_TABLES=1): the correspondence search, the on-demand target
// covariances and the fitness score read the source's sorted points only — its level tables serve its own k-NN,
// on lane 1, and the covariances that lane is still computing are only needed by the first kernel that reads them
// (k_linearize), which calls join_side() itself.  side_pending stays set.  The host has already run the build's
// last phase (reg_drain), so a sort redone after a grid-geometry miss is the one the event stands for.
static int join_side_build(rgc_ctx* c, const Cloud& cl) {
  static const bool tables = std::getenv("RGC_JOIN_TABLES") != nullptr;
  cudaEvent_t ev = (!tables && cl.ev[5]) ? cl.ev[5] : cl.ev[1];
  if (c->side_pending && cl.valid && cl.lane_built == 1 && ev) CK(c, cudaStreamWaitEvent(c->stream, ev, 0));
  return RGC_OK;
}

static void cloud_release(rgc_ctx* c, Cloud& cl) {
  if (cl.job) {  // a build in flight: its scratch goes back to the pool (stream-ordered reuse), the rest below
    if (cl.job->spec_tables) cudaStreamSynchronize(c->aux[cl.job->lane]);  // (error paths only) the forked count must not outlive its scratch
    c->put_hslot(cl.job->h_slot);
    c->put_event(cl.job->ready);
    c->put_event(cl.job->cleared);
    cl.job.reset();
  }
  c->put(cl.sorted);
  c->put(cl.inv);
  c->put(cl.tables);
  c->put(cl.cov);
  c->put(cl.cov_state);
  c->put(cl.d_off);
  c->put(cl.d_tiles);
  for (int i = 0; i < 6; i++) c->put_event(cl.ev[i]);
  cl = Cloud();
}

// a cloud that lives for one call (rgc_knn, debug hooks): released on every exit path
struct TmpCloud {
  rgc_ctx* c;
  Cloud cl;
  explicit TmpCloud(rgc_ctx* c_) : c(c_) {}
  ~TmpCloud() { cloud_release(c, cl); }
};

// stable LSD radix sort of (64-bit key, 32-bit value) pairs on the ctx stream, one launch per 8-bit digit
// (k_rs_onesweep).  `scratch`: rs_scratch_words(n, passes) words.  `ghist_ready`: the caller zeroed the scratch
// and already counted the digit histograms (k_keys_hist).  `gather` (nullable): the last pass writes the points
// in sorted order + the inverse permutation instead of the value array.  On return *kout / *vout point at the
// buffers with the sorted data.
struct SortGather {
  const float4* pts;
  float4* sorted;
  int* inv;
};
static inline int sort_passes(int key_bits) { return std::max(1, div_up(key_bits, 8)); }
static int radix_sort_pairs(rgc_ctx* c, uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint32_t* scratch, int n, int key_bits,
                            uint64_t** kout, uint32_t** vout, bool ghist_ready = false, const SortGather* gather = nullptr) {
  const int nblk = div_up(n, RS_TILE);
  const int passes = sort_passes(key_bits);
  if (passes > RS_MAX_PASSES) FAIL(c, RGC_ERR_UNSUPPORTED, "sort key wider than 64 bits");
  if (!ghist_ready) {
    CK(c, cudaMemsetAsync(scratch, 0, sizeof(uint32_t) * rs_scratch_words(n, passes), c->stream));
    k_rs_ghist<<<std::min(div_up(n, 256 * 8), 148 * 4), 256, 0, c->stream>>>(keys_a, n, passes, scratch);
    CKL(c);
  }
  uint64_t *kin = keys_a, *ko = keys_b;
  uint32_t *vin = vals_a, *vo = vals_b;
  for (int p = 0; p < passes; p++) {
    if (gather && p == passes - 1)
      launch_pdl(c, k_rs_onesweep<true>, dim3(nblk), dim3(256), 0, c->stream, kin, vin, ko, vo, scratch, p, n, nblk, gather->pts, gather->sorted, gather->inv);
    else
      launch_pdl(c, k_rs_onesweep<false>, dim3(nblk), dim3(256), 0, c->stream, kin, vin, ko, vo, scratch, p, n, nblk, nullptr, nullptr, nullptr);
    CKL(c);
    std::swap(kin, ko);
    std::swap(vin, vo);
  }
  *kout = kin;
  *vout = vin;
  return RGC_OK;
}

static int build_keys(rgc_ctx* c, Cloud& cl, const unsigned char* raw, size_t stride);
static int build_passes(rgc_ctx* c, Cloud& cl);
static int build_tables_spec(rgc_ctx* c, Cloud& cl);

// level tables of `slots[l]` entries each, laid out one after the other in cl.tables
static void table_layout(Cloud& cl, const size_t* slots, TableSet& ts) {
  GridView& v = cl.view;
  ts.nlevels = v.nlevels;
  size_t off = 0;
  for (int l = 0; l < kMaxLevels; l++) {
    if (l < v.nlevels) {
      ts.table[l] = cl.tables + off;
      ts.mask[l] = (uint32_t)(slots[l] - 1);
      int lg = 0;
      while (((size_t)1 << lg) < slots[l]) lg++;
      ts.shift[l] = (uint32_t)(64 - lg);
      off += slots[l];
    } else {
      ts.table[l] = nullptr;
      ts.mask[l] = 0;
      ts.shift[l] = 63;
    }
    v.table[l] = ts.table[l];
    v.mask[l] = ts.mask[l];
    v.shift[l] = ts.shift[l];
  }
}
static size_t table_slots_for(uint32_t cells) {
  size_t s = 8;
  while (s < 2 * (size_t)cells) s <<= 1;
  return s;
}

// upload (or adopt a device pointer), Morton-sort, build the level tables — in three phases separated by
// the two host waits (see BuildJob).  All phases of a cloud run on the lane that was current in phase 1.
// `offsets` (nullable, n_clouds + 1 entries): the input is the concatenation of n_clouds clouds that
// become one multi-cloud grid (rgc_grid.cuh: CloudRange)
static int build_phase1(rgc_ctx* c, Cloud& cl, const void* points, size_t n_sz, size_t stride, bool on_device, uint64_t key, float cell, const int* offsets,
                        int n_clouds, bool defer_passes = false) {
  cloud_release(c, cl);
  if (n_sz == 0 || points == nullptr) FAIL(c, RGC_ERR_INVALID, "empty point cloud");
  if (n_sz > 0x7fffffff / 32) FAIL(c, RGC_ERR_UNSUPPORTED, "point cloud too large for 32-bit indexing");
  if (stride < 12 || stride % 4) FAIL(c, RGC_ERR_INVALID, "point stride must be a multiple of 4 and >= 12 bytes");
  const int n = (int)n_sz;
  cudaStream_t st = c->stream;
  for (int i = 0; i < 6; i++)
    if (!(cl.ev[i] = c->get_event())) FAIL(c, RGC_ERR_CUDA, "cudaEventCreate failed");
  CK(c, cudaEventRecord(cl.ev[0], st));
  c->mark("build: begin");
  cl.job.reset(new BuildJob(c));
  BuildJob& j = *cl.job;
  j.stage = 1;
  j.lane = c->lane;
  j.n = n;
  j.key = key;
  j.cell = cell;
  j.h_slot = c->get_hslot();
  j.d_slot = rgc_ctx::hslot_device(j.h_slot);
  j.ready = c->get_event();
  if (!j.h_slot || !j.d_slot || !j.ready) FAIL(c, RGC_ERR_NOMEM, "pinned slot / event allocation failed (build)");
  if (offsets) {
    while ((1 << j.cloud_bits) < n_clouds) j.cloud_bits++;
    j.n_clouds = n_clouds;
    cl.n_clouds = n_clouds;
    cl.h_off.assign(offsets, offsets + n_clouds + 1);
    cl.d_off = (int*)c->get(sizeof(int) * (size_t)(n_clouds + 1));
    if (!cl.d_off) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (cloud offsets)");
    CK(c, cudaMemcpyAsync(cl.d_off, offsets, sizeof(int) * (size_t)(n_clouds + 1), cudaMemcpyHostToDevice, st));
  }
  const unsigned char* d_raw = (const unsigned char*)points;
  if (!on_device) {
    void* staging = j.tmp.get(n_sz * stride);
    if (!staging) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (staging)");
    CK(c, cudaMemcpyAsync(staging, points, n_sz * stride, cudaMemcpyHostToDevice, st));
    d_raw = (const unsigned char*)staging;
  }
  j.orig = (float4*)j.tmp.get(sizeof(float4) * n_sz);
  j.keys_a = (uint64_t*)j.tmp.get(8 * n_sz);
  j.keys_b = (uint64_t*)j.tmp.get(8 * n_sz);
  j.vals_a = (uint32_t*)j.tmp.get(4 * n_sz);
  j.vals_b = (uint32_t*)j.tmp.get(4 * n_sz);
  j.hist = (uint32_t*)j.tmp.get(4 * rs_scratch_words(n, RS_MAX_PASSES));
  j.d_counts = (uint32_t*)j.tmp.get(4 * (kMaxLevels + 1));
  cl.sorted = (float4*)c->get(sizeof(float4) * n_sz);
  cl.inv = (int*)c->get(sizeof(int) * n_sz);
  if (!cl.inv || !j.orig || !j.keys_a || !j.keys_b || !j.vals_a || !j.vals_b || !j.hist || !j.d_counts || !cl.sorted)
    FAIL(c, RGC_ERR_NOMEM, "device allocation failed (build)");
  j.max_bits = std::min(kMaxBits, (56 - j.cloud_bits) / 3);  // key = cloud id | 3 * nbits Morton bits <= 56 bits
  // Speculative geometry: successive clouds of a stream (the sweeps / the submaps of a SLAM front end) need the
  // same grid, so the keys are generated with the previous cloud's (nbits, s0) in the SAME kernel that ingests
  // the points and reduces the bounding box, and the sort follows without the host waiting for the box.
  // build_phase3 checks the box against that geometry when it arrives with the level counts (one host wait
  // per cloud instead of two) and redoes phase 2 the slow way if it does not fit.  A larger grid than needed
  // is still a valid grid with the same Morton order (rgc_grid.cuh), so the results do not depend on the hint.
  const rgc_ctx::GeomHint& h = c->geom_hint[c->lane];
  if (c->spec_build && h.valid && h.cell == cell && h.cloud_bits == j.cloud_bits && h.nbits <= j.max_bits) {
    j.spec = true;
    c->spec_builds++;
    grid_set(cl.view, h.nbits, h.s0);
    cl.view.n = n;
    // (only for a cloud of about the previous one's size: its tables were sized for 25 - 50 % load, and a table that
    // runs full makes every further insert probe all of it — a 19.6 M-point multi-cloud target grid filled into the
    // tables of the 4.4 M-point source grid built just before it on the same lane did not finish in minutes)
    if (c->spec_tables && h.have_slots && (long long)n * 4 <= (long long)h.slots_n * 5) {
      // level tables of the previous cloud's sizes, cleared on the aux stream while the keys are made and sorted
      size_t total = 0;
      for (int l = 0; l < cl.view.nlevels; l++) total += (j.spec_slots[l] = h.slots[l]);
      cl.tables = (GridSlot*)c->get(total * sizeof(GridSlot));
      j.cleared = c->get_event();
      if (!cl.tables || !j.cleared) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (tables)");
      cudaStream_t aux = c->aux[c->lane];
      CK(c, cudaStreamWaitEvent(aux, cl.ev[0], 0));  // the pool block's previous users are ordered before ev[0] on this lane
      CK(c, cudaMemsetAsync(cl.tables, 0xff, total * sizeof(GridSlot), aux));
      CK(c, cudaEventRecord(j.cleared, aux));
      j.spec_tables = true;
      c->spec_table_builds++;
    }
    TRY(build_keys(c, cl, d_raw, stride));
    // main-lane (target) clouds of a deferred build stop here: their sort passes are issued by the drain AFTER the
    // side lane's (the source's whole chain is short but the host needs ~40 us to issue a sort: issued first, the
    // target's passes kept the source from even starting for 70 us)
    if (defer_passes) return RGC_OK;
    return build_passes(c, cl);
  }
  k_ingest<<<kBboxBlocks, 256, 0, st>>>(d_raw, stride, n, j.orig, j.d_slot);
  CKL(c);
  CK(c, cudaEventRecord(j.ready, st));
  return RGC_OK;
}

// bounding-box partials of a build (pinned copy) -> box; false if any coordinate was not finite
static bool build_bbox(const BuildJob& j, float mn[3], float mx[3]) {
  bool finite = true;  // a block that saw NaN / inf writes NaN partials (k_ingest)
  for (int a = 0; a < 3; a++) {
    mn[a] = INFINITY;
    mx[a] = -INFINITY;
  }
  for (int b = 0; b < kBboxBlocks; b++)
    for (int a = 0; a < 3; a++) {
      finite = finite && !std::isnan(j.h_slot[b * 6 + a]) && !std::isnan(j.h_slot[b * 6 + 3 + a]);
      mn[a] = std::min(mn[a], j.h_slot[b * 6 + a]);
      mx[a] = std::max(mx[a], j.h_slot[b * 6 + 3 + a]);
    }
  for (int a = 0; a < 3; a++) finite = finite && std::isfinite(mn[a]) && std::isfinite(mx[a]);
  return finite;
}

// Morton keys + digit histograms of all sort passes (+ ingest and bounding-box partials when `raw` is given: the
// speculative path).  cl.view holds the geometry (nbits, s0) to use.
static int build_keys(rgc_ctx* c, Cloud& cl, const unsigned char* raw, size_t stride) {
  BuildJob& j = *cl.job;
  cudaStream_t st = c->stream;
  const int n = j.n;
  const GridView& v = cl.view;
  const int passes = sort_passes(3 * v.nbits + j.cloud_bits);
  const GridGeom geom{v.inv_s0, v.bias, v.nbits};
  CK(c, cudaMemsetAsync(j.hist, 0, sizeof(uint32_t) * rs_scratch_words(n, passes), st));
  if (raw)
    launch_pdl(c, k_keys_hist<true>, dim3(kBboxBlocks), dim3(256), 0, st, raw, stride, n, j.orig, j.d_slot, geom, cl.d_off, j.n_clouds, passes, j.keys_a, j.vals_a, j.hist);
  else
    launch_pdl(c, k_keys_hist<false>, dim3(kBboxBlocks), dim3(256), 0, st, nullptr, 0, n, j.orig, nullptr, geom, cl.d_off, j.n_clouds, passes, j.keys_a, j.vals_a, j.hist);
  CKL(c);
  c->mark("build: keys + histograms");
  j.stage = 3;
  return RGC_OK;
}
// radix sort with the gather fused into the last pass -> per-level cell counts, published in the mapped slot -> `ready`
static int build_passes(rgc_ctx* c, Cloud& cl) {
  BuildJob& j = *cl.job;
  cudaStream_t st = c->stream;
  const int n = j.n;
  const GridView& v = cl.view;
  j.kin = j.keys_a;
  j.vin = j.vals_a;
  const SortGather gather{j.orig, cl.sorted, cl.inv};
  TRY(radix_sort_pairs(c, j.keys_a, j.keys_b, j.vals_a, j.vals_b, j.hist, n, 3 * v.nbits + j.cloud_bits, &j.kin, &j.vin, true, &gather));
  CK(c, cudaEventRecord(cl.ev[5], st));
  c->mark("build: radix sort + gather");
  // level counts, then the sort's error word (a look-back that gave up: cannot happen, but must not pass silently),
  // written into the mapped slot by the kernel's last block
  cudaStream_t cst = st;
  if (j.spec_tables) {
    // the tables are filled right away, with the sizes of the lane's previous cloud; the counts that say whether those
    // sizes hold are taken beside it on the aux stream (build_phase3 reads them)
    TRY(build_tables_spec(c, cl));
    cst = c->aux[j.lane];
    CK(c, cudaStreamWaitEvent(cst, cl.ev[5], 0));
  }
  CK(c, cudaMemsetAsync(j.d_counts, 0, 4 * (kMaxLevels + 1), cst));
  launch_pdl(c, k_count_cells, dim3(div_up(n, 256)), dim3(256), 0, cst, j.kin, n, v.nlevels, j.d_counts, j.hist + 8, reinterpret_cast<uint32_t*>(j.d_slot + 6 * kBboxBlocks));
  CKL(c);
  CK(c, cudaEventRecord(j.ready, cst));
  if (!j.spec_tables) c->mark("build: cell counts");
  j.stage = 2;
  return RGC_OK;
}
// fill the level tables with the sizes of the lane's previous cloud, right behind the sort (BuildJob::spec_tables)
static int build_tables_spec(rgc_ctx* c, Cloud& cl) {
  BuildJob& j = *cl.job;
  cudaStream_t st = c->stream;
  TableSet ts;
  table_layout(cl, j.spec_slots, ts);
  CK(c, cudaStreamWaitEvent(st, j.cleared, 0));
  launch_pdl(c, k_build_tables, dim3(div_up(j.n, 256), cl.view.nlevels), dim3(256), 0, st, j.kin, j.n, ts);
  CKL(c);
  CK(c, cudaEventRecord(cl.ev[1], st));
  c->mark("build: tables (speculative sizes)");
  return RGC_OK;
}

static int build_sort(rgc_ctx* c, Cloud& cl, const unsigned char* raw, size_t stride) {
  TRY(build_keys(c, cl, raw, stride));
  return build_passes(c, cl);
}

// bounding box on the host -> grid geometry, then the sort (non-speculative path)
static int build_phase2(rgc_ctx* c, Cloud& cl) {
  BuildJob& j = *cl.job;
  CK(c, cudaEventSynchronize(j.ready));
  float mn[3], mx[3];
  if (!build_bbox(j, mn, mx)) FAIL(c, RGC_ERR_INVALID, "point cloud contains non-finite coordinates");
  for (int a = 0; a < 3; a++) {
    cl.bb_min[a] = mn[a];
    cl.bb_max[a] = mx[a];
  }
  GridView& v = cl.view;
  v.n = j.n;
  grid_geometry(mn, mx, j.cell, v, j.max_bits);
  j.spec = false;
  return build_sort(c, cl, nullptr, 0);
}

// level counts on the host -> table sizes, table build (with the child masks); the cloud becomes valid
static int build_phase3(rgc_ctx* c, Cloud& cl) {
  BuildJob& j = *cl.job;
  CK(c, cudaEventSynchronize(j.ready));
  cudaStream_t st = c->stream;
  const int n = j.n;
  GridView& v = cl.view;
  if (j.spec) {
    // the bounding box arrived with the counts: does the geometry the keys were made with hold it?
    float mn[3], mx[3];
    if (!build_bbox(j, mn, mx)) FAIL(c, RGC_ERR_INVALID, "point cloud contains non-finite coordinates");
    for (int a = 0; a < 3; a++) {
      cl.bb_min[a] = mn[a];
      cl.bb_max[a] = mx[a];
    }
    int nbits;
    float s0;
    grid_bits(mn, mx, j.cell, j.max_bits, nbits, s0);
    rgc_ctx::GeomHint& h = c->geom_hint[j.lane];
    if (s0 != v.s0 || nbits > v.nbits) {  // it does not: sort again with the right geometry (j.orig is intact)
      c->spec_misses++;
      if (j.spec_tables) {  // the tables filled from those keys go too (stream-ordered reuse of the block on this lane)
        c->put(cl.tables);
        cl.tables = nullptr;
        j.spec_tables = false;
      }
      grid_geometry(mn, mx, j.cell, v, j.max_bits);
      v.n = n;
      j.spec = false;
      TRY(build_sort(c, cl, nullptr, 0));
      CK(c, cudaEventSynchronize(j.ready));
    } else {
      v.margin = grid_margin(mn, mx);
      if (nbits + 1 < v.nbits) h.nbits = nbits + 1;  // much smaller clouds now: shrink the hint (with one bit of hysteresis)
    }
  }
  {
    rgc_ctx::GeomHint& h = c->geom_hint[j.lane];
    if (!j.spec || !h.valid) h = rgc_ctx::GeomHint{true, v.nbits, v.s0, j.cell, j.cloud_bits};
  }
  const uint32_t* h_counts = reinterpret_cast<const uint32_t*>(j.h_slot + 6 * kBboxBlocks);
  if (h_counts[kMaxLevels] != 0) FAIL(c, RGC_ERR_CUDA, "radix sort look-back timed out");
  size_t total_slots = 0;
  size_t slots[kMaxLevels];
  for (int l = 0; l < v.nlevels; l++) {
    slots[l] = table_slots_for(h_counts[l]);
    total_slots += slots[l];
  }
  {  // what the next cloud of this lane will try
    rgc_ctx::GeomHint& h = c->geom_hint[j.lane];
    h.have_slots = h.valid && h.nbits == v.nbits && v.nlevels <= 20;
    for (int l = 0; h.have_slots && l < v.nlevels; l++) h.slots[l] = (uint32_t)slots[l];
    h.slots_n = n;
  }
  bool tables_done = false;
  if (j.spec_tables) {
    // tables already filled (build_passes) with the previous cloud's sizes: they stand if every level stayed under
    // 70 % load (sized from its own counts a table is 25 - 50 % full); the search results do not depend on the sizes
    tables_done = true;
    for (int l = 0; l < v.nlevels; l++) tables_done = tables_done && 10 * (size_t)h_counts[l] <= 7 * j.spec_slots[l];
    if (!tables_done) {
      c->spec_table_misses++;
      c->put(cl.tables);
      cl.tables = nullptr;
    }
  }
  if (!tables_done) {
    cl.tables = (GridSlot*)c->get(total_slots * sizeof(GridSlot));
    if (!cl.tables) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (tables)");
    CK(c, cudaMemsetAsync(cl.tables, 0xff, total_slots * sizeof(GridSlot), st));
    TableSet ts;
    table_layout(cl, slots, ts);
    c->mark("build: host wait over, tables cleared");
    launch_pdl(c, k_build_tables, dim3(div_up(n, 256), v.nlevels), dim3(256), 0, st, j.kin, n, ts);
    CKL(c);
    c->mark("build: tables");
    CK(c, cudaEventRecord(cl.ev[1], st));
  }
  v.pts = reinterpret_cast<const F4*>(cl.sorted);
  v.inv = cl.inv;
  cl.n = n;
  cl.key = j.key;
  cl.valid = true;
  cl.lane_built = c->lane;
  cl.build_timed = true;
  c->put_hslot(j.h_slot);
  c->put_event(j.ready);
  c->put_event(j.cleared);
  cl.job.reset();  // scratch back to the pool (the kernels above are ordered before any reuse on this lane)
  return RGC_OK;
}

// the whole build, synchronously, on the current lane
static int cloud_build(rgc_ctx* c, Cloud& cl, const void* points, size_t n_sz, size_t stride, bool on_device, uint64_t key, float cell,
                       const int* offsets = nullptr, int n_clouds = 0) {
  int rc = build_phase1(c, cl, points, n_sz, stride, on_device, key, cell, offsets, n_clouds);
  if (rc == RGC_OK && cl.job->stage == 1) rc = build_phase2(c, cl);
  if (rc == RGC_OK && cl.job->stage == 3) rc = build_passes(c, cl);
  if (rc == RGC_OK) rc = build_phase3(c, cl);
  if (rc != RGC_OK) cloud_release(c, cl);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// Pre-step: host cloud -> [de-skew] -> [pcl::VoxelGrid centroids] as a pooled device
// buffer of float4 (x, y, z, intensity) on the current lane.  *d_out must be given back with c->put().
static int pre_filter(rgc_ctx* c, const void* points, size_t n_sz, size_t stride, size_t inten_off, float leaf, const double* q_wxyz, const double* t3,
                      float scan_period, float4** d_out, int* n_out, int* passthrough) {
  *d_out = nullptr;
  *n_out = 0;
  if (passthrough) *passthrough = 0;
  if (n_sz == 0 || points == nullptr) FAIL(c, RGC_ERR_INVALID, "empty point cloud");
  if (n_sz > 0x7fffffff / 32) FAIL(c, RGC_ERR_UNSUPPORTED, "point cloud too large for 32-bit indexing");
  if (stride < 12 || stride % 4) FAIL(c, RGC_ERR_INVALID, "point stride must be a multiple of 4 and >= 12 bytes");
  if (inten_off != kNoIntensity && (inten_off % 4 || inten_off + 4 > stride)) FAIL(c, RGC_ERR_INVALID, "intensity offset outside the point record");
  if (q_wxyz && (!t3 || !(scan_period > 0.f))) FAIL(c, RGC_ERR_INVALID, "de-skew needs q, t and a positive scan period");
  if (q_wxyz && inten_off == kNoIntensity) FAIL(c, RGC_ERR_INVALID, "de-skew needs the intensity channel (ring + relative time)");
  const int n = (int)n_sz;
  cudaStream_t st = c->stream;
  DeskewParams D{};
  if (q_wxyz) {
    const double qw = q_wxyz[0], qx = q_wxyz[1], qy = q_wxyz[2], qz = q_wxyz[3];
    const double n2 = ((qx * qx + qy * qy) + qz * qz) + qw * qw;  // Eigen squaredNorm, coefficient order x y z w
    D.enabled = 1;
    if (n2 > 0.0) {
      D.iw = qw / n2;
      D.ix = -qx / n2;
      D.iy = -qy / n2;
      D.iz = -qz / n2;
    }
    D.tx = t3[0];
    D.ty = t3[1];
    D.tz = t3[2];
    D.scan_period = scan_period;
  }
  Scratch tmp(c);  // returned to the pool on every exit path; the block handed to the caller is release()d
  void* staging = tmp.get(n_sz * stride);
  float4* pts = (float4*)tmp.get(sizeof(float4) * n_sz);
  float* d_bbox = (float*)tmp.get(sizeof(float) * 6 * kBboxBlocks);
  if (!staging || !pts || !d_bbox) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (pre-step)");
  CK(c, cudaMemcpyAsync(staging, points, n_sz * stride, cudaMemcpyHostToDevice, st));
  k_pre_ingest<<<kBboxBlocks, 256, 0, st>>>((const unsigned char*)staging, stride, inten_off, n, D, pts, d_bbox);
  CKL(c);
  if (!(leaf > 0.f)) {  // de-skew only
    *d_out = (float4*)tmp.release(pts);
    *n_out = n;
    return RGC_OK;
  }
  CK(c, cudaMemcpyAsync(c->h_bbox, d_bbox, sizeof(float) * 6 * kBboxBlocks, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int b = 0; b < kBboxBlocks; b++)
    for (int a = 0; a < 3; a++) {
      // pcl::VoxelGrid skips non-finite points of a non-dense cloud; this interface requires finite input
      // (include/rgc_preprocess.h) and says so instead of silently mis-binning them
      if (std::isnan(c->h_bbox[b * 6 + a]) || std::isnan(c->h_bbox[b * 6 + 3 + a])) FAIL(c, RGC_ERR_INVALID, "point cloud contains non-finite coordinates");
      mn[a] = std::min(mn[a], c->h_bbox[b * 6 + a]);
      mx[a] = std::max(mx[a], c->h_bbox[b * 6 + 3 + a]);
    }
  // pcl::VoxelGrid::applyFilter (filters/impl/voxel_grid.hpp): float arithmetic throughout
  VgGeom vg;
  vg.inv_leaf = 1.0f / leaf;
  long long dd[3];
  for (int a = 0; a < 3; a++) dd[a] = (long long)((mx[a] - mn[a]) * vg.inv_leaf) + 1;
  if (dd[0] * dd[1] * dd[2] > (long long)INT_MAX) {
    // "Leaf size is too small for the input dataset. Integer indices would overflow.": PCL returns the input
    if (passthrough) *passthrough = 1;
    *d_out = (float4*)tmp.release(pts);
    *n_out = n;
    return RGC_OK;
  }
  int div_b[3];
  for (int a = 0; a < 3; a++) {
    vg.min_b[a] = (int)std::floor(mn[a] * vg.inv_leaf);
    div_b[a] = (int)std::floor(mx[a] * vg.inv_leaf) - vg.min_b[a] + 1;
  }
  vg.mul1 = div_b[0];
  vg.mul2 = div_b[0] * div_b[1];
  int key_bits = 1;
  while (key_bits < 32 && (1ll << key_bits) < (long long)div_b[0] * div_b[1] * div_b[2]) key_bits++;
  const int nblk_rs = div_up(n, RS_TILE), nblk_sc = div_up(n, 256 * kScanItems);
  uint64_t* keys_a = (uint64_t*)tmp.get(8 * n_sz);
  uint64_t* keys_b = (uint64_t*)tmp.get(8 * n_sz);
  uint32_t* vals_a = (uint32_t*)tmp.get(4 * n_sz);
  uint32_t* vals_b = (uint32_t*)tmp.get(4 * n_sz);
  uint32_t* hist = (uint32_t*)tmp.get(4 * rs_scratch_words(n, RS_MAX_PASSES));
  unsigned int* blk = (unsigned int*)tmp.get(4 * ((size_t)nblk_sc + 1));
  int* heads = (int*)tmp.get(4 * n_sz);
  if (!keys_a || !keys_b || !vals_a || !vals_b || !hist || !blk || !heads) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (voxel filter)");
  k_vg_keys<<<div_up(n, 256), 256, 0, st>>>(pts, n, vg, keys_a, vals_a);
  CKL(c);
  uint64_t* ks = keys_a;
  uint32_t* vs = vals_a;
  TRY(radix_sort_pairs(c, keys_a, keys_b, vals_a, vals_b, hist, n, key_bits, &ks, &vs));
  k_vg_head_count<<<nblk_sc, 256, 0, st>>>(ks, n, blk);
  CKL(c);
  k_vg_scan_blocks<<<1, 1024, 0, st>>>(blk, nblk_sc);
  CKL(c);
  k_vg_heads<<<nblk_sc, 256, 0, st>>>(ks, n, blk, heads);
  CKL(c);
  CK(c, cudaMemcpyAsync(c->h_counts, blk + nblk_sc, 4, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  const int nv = (int)c->h_counts[0];
  float4* out = (float4*)c->get(sizeof(float4) * (size_t)nv);
  if (!out) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (voxel filter output)");
  k_vg_centroid<<<div_up(nv, 128), 128, 0, st>>>(pts, vs, heads, nv, n, out);
  c->launches++;
  if (cudaGetLastError() != cudaSuccess) {
    c->put(out);
    FAIL(c, RGC_ERR_CUDA, "kernel launch: k_vg_centroid");
  }
  *d_out = out;
  *n_out = nv;
  return RGC_OK;
}

// Largest k (setCorrespondenceRandomness / rgc_knn).  The fast paths — the warp-per-query kernel that finishes deferred
// tiles and serves the on-demand target covariances, the unrolled covariance kernels, the batched registration —
// hold k <= 32; above that the tile kernel alone does the self-kNN (its per-lane heaps live in shared memory:
// 36 KB per warp at k = 128), the covariances come from the generic kernel and the target's are computed eagerly.
constexpr int kMaxK = 128;
#ifndef RGC_KNN_ALLWARP_MAX  // largest single cloud whose self-kNN runs entirely in the warp-per-query kernel (0: never)
#define RGC_KNN_ALLWARP_MAX 48000
#endif

template <bool SELF>
static int launch_knn(rgc_ctx* c, const GridView& v, const float4* queries, int m, int k, int* idx, float* d2) {
  if (k > kMaxK) FAIL(c, RGC_ERR_UNSUPPORTED, "k > 128 is not supported");
  const int spread = query_spread(m);
  const int grid = div_up(m * spread, kThreads);
  const size_t smem = (size_t)k * kThreads * 8;  // per-thread max-heap of k (d2, position) pairs
  if (smem > 48 * 1024) CK(c, cudaFuncSetAttribute(k_knn<SELF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_knn<SELF><<<grid, kThreads, smem, c->stream>>>(v, queries, m, k, spread, idx, d2);
  CKL(c);
  return RGC_OK;
}

// self-kNN of a whole cloud: the warp-cooperative tile kernel (RGC_KNN_THREAD=1 selects the
// thread-per-query kernel instead, for A/B profiling)
static int launch_knn_self(rgc_ctx* c, const GridView& v, int n, int k, int* nbr, const TileDesc* tiles = nullptr, int n_tiles = 0, int n_clouds = 1) {
  // the warp-cooperative tile kernel; RGC_KNN_THREAD=1 selects the thread-per-query kernel (A/B, single clouds only)
  static const bool per_thread = std::getenv("RGC_KNN_THREAD") != nullptr;
  if (per_thread && !tiles) return launch_knn<true>(c, v, nullptr, n, k, nbr, nullptr);
  if (k > kMaxK) FAIL(c, RGC_ERR_UNSUPPORTED, "k_correspondences > 128 is not supported");
  const size_t per_warp = sizeof(float4) * KT_CAND + sizeof(TileNode) * KT_STACK + (size_t)(k + KT_PEND) * 32 * 8;
  const size_t smem = per_warp * KT_WARPS;
  CK(c, cudaFuncSetAttribute(k_knn_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_cloud = n / std::max(n_clouds, 1);  // typical size of ONE cloud (a multi-cloud grid holds many)
  // more seeds pay off on a single sweep (sparse rings: tighter start bounds), fewer on dense maps (sweep in profiles/README.md)
  const int n_seeds = n_cloud < 100000 ? KT_SEEDS : (KT_SEEDS > 64 ? 64 : KT_SEEDS);
  // tiles that gather more than `defer` candidates are finished by k_knn_warp (one warp per query):
  // they are the sparse-region tiles that used to form a 40 % tail of this launch (profiles/README.md)
  // (the tail is one slow tile long, ~0.3 ms, whatever n is, while deferring costs ~8 % extra work:
  // it pays below a few million points; measured 13.1 vs 14.4 ms at 8 M points, 1.33 vs 0.98 ms at 500 k)
  // single sweeps (sparse rings, < 1 wave of tiles): an early hand-over shortens the launch, 0.23 -> 0.14 ms
  // on a 22k-point sweep; dense maps keep their tiles longer (600: 2.7 vs 1.7 ms on the 500k submap at 150)
  const int auto_defer = n_cloud < 100000 ? 150 : 600;
  const int want_defer = c->knn_defer < 0 ? auto_defer : c->knn_defer;
  const int defer = (want_defer > 0 && n_cloud < 2000000 && k <= 32) ? want_defer : INT_MAX;  // k_knn_warp keeps the k best in one warp's registers
  const int ntiles = tiles ? n_tiles : div_up(n, 32);
  // a single sweep-sized cloud skips the tile kernel: one warp per query for every point.  With the four-node walk and
  // the batch merge of round 2 the warp kernel does the 22k-point sweep in 0.077 ms; the tile kernel + the warp
  // kernel for its deferred tiles took 0.165 ms (a sweep's sparse rings make most tiles gather far more candidates
  // than 32 queries can share).  Dense maps keep the tile kernel (2.6x fewer instructions per query there).
  // (an explicit deferral threshold — rgc_debug_set_knn_defer / RGC_KNN_DEFER, the tests' way of forcing either kernel — keeps the tile path)
  if (!tiles && k <= 32 && n <= RGC_KNN_ALLWARP_MAX && c->knn_defer < 0) {
    launch_pdl(c, k_knn_warp, dim3(std::min(div_up(n, KW_WARPS), 148 * RGC_KW_MINB)), dim3(KW_WARPS * 32), 0, c->stream, v, n, k, nullptr, nullptr, nullptr, 0, nbr, nullptr, nullptr, 0);
    CKL(c);
    return RGC_OK;
  }
  Scratch tmp(c);
  int* dq = (int*)tmp.get(sizeof(int) * (size_t)(ntiles + 1));  // [0] = count, [1..] = tile ids
  if (!dq) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (knn defer list)");
  CK(c, cudaMemsetAsync(dq, 0, sizeof(int), c->stream));
  launch_pdl(c, k_knn_tile, dim3(div_up(ntiles, KT_WARPS)), dim3(KT_WARPS * 32), smem, c->stream, v, n, k, n_seeds, defer, dq, dq + 1, nbr, tiles, ntiles);
  CKL(c);
  if (defer != INT_MAX) {
    launch_pdl(c, k_knn_warp, dim3(std::min(div_up(n, KW_WARPS), 148 * RGC_KW_MINB)), dim3(KW_WARPS * 32), 0, c->stream, v, n, k, dq, dq + 1, nullptr, 0, nbr, tiles, nullptr, 0);
    CKL(c);
  }
  return RGC_OK;
}

// tile list of a multi-cloud grid for the self-kNN: tiles of 32 consecutive sorted positions, none straddling two clouds
static int cloud_tiles(rgc_ctx* c, Cloud& cl) {
  if (cl.n_clouds <= 0) return RGC_OK;
  std::vector<TileDesc> t;
  t.reserve((size_t)cl.n / 32 + (size_t)cl.n_clouds);
  for (int q = 0; q < cl.n_clouds; q++) {
    const int lo = cl.h_off[q], hi = cl.h_off[q + 1];
    const uint64_t prefix = (uint64_t)q << (3 * cl.view.nbits);
    for (int first = lo; first < hi; first += 32) t.push_back(TileDesc{first, lo, hi, 0, prefix});
  }
  c->put(cl.d_tiles);
  cl.n_tiles = (int)t.size();
  cl.d_tiles = (TileDesc*)c->get(sizeof(TileDesc) * t.size());
  if (!cl.d_tiles) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (tile list)");
  // pageable source: the copy is staged before the call returns, so the vector may go out of scope
  CK(c, cudaMemcpyAsync(cl.d_tiles, t.data(), sizeof(TileDesc) * t.size(), cudaMemcpyHostToDevice, c->stream));
  return RGC_OK;
}

// k_covariance over `grid_threads` threads (whole cloud: thread = sorted point; on-demand: thread = entry of `qlist`).
// `full`: k == 20 and every neighbour slot is filled (the predicate-free instantiation).
static int launch_covariance(rgc_ctx* c, cudaStream_t st, const float4* pts, const int* nbr, int n_stride, int grid_threads, int k, bool full, int method,
                             double* cov, const int* qlist, const int* qcount) {
  const int grid = div_up(grid_threads, kThreads);
#if RGC_ASYNC_STAGE
  const size_t smem20 = sizeof(float4) * 20 * kThreads, smem32 = sizeof(float4) * 32 * kThreads;  // staged neighbour points
  static bool attr_set = false;
  if (!attr_set) {
    CK(c, cudaFuncSetAttribute(k_covariance<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
    attr_set = true;
  }
#else
  const size_t smem20 = 0, smem32 = 0;
#endif
  if (k > 32)
    k_covariance_any<<<grid, kThreads, 0, st>>>(pts, nbr, n_stride, grid_threads, k, method, cov);
  else if (k == 20 && full)
    launch_pdl(c, k_covariance<20, true>, dim3(grid), dim3(kThreads), smem20, st, pts, nbr, n_stride, k, method, cov, qlist, qcount);
  else if (k <= 20)
    launch_pdl(c, k_covariance<20, false>, dim3(grid), dim3(kThreads), smem20, st, pts, nbr, n_stride, k, method, cov, qlist, qcount);
  else
    launch_pdl(c, k_covariance<32, false>, dim3(grid), dim3(kThreads), smem32, st, pts, nbr, n_stride, k, method, cov, qlist, qcount);
  CKL(c);
  return RGC_OK;
}

// FastGICP::calculate_covariances (fast_gicp_impl.hpp:241-299)
static int cloud_covariances(rgc_ctx* c, Cloud& cl, int k, int method, bool speculative = false) {
  // covariances computed ahead of time (set_input) are redone if the parameters changed before the
  // first align; once an align has used them they stay, like the reference's (recomputed only when
  // the cloud changes)
  if (cl.has_cov && !(cl.cov_speculative && (cl.cov_k != k || cl.cov_method != method))) {
    if (!speculative) cl.cov_speculative = false;
    return RGC_OK;
  }
  if (k < 1) FAIL(c, RGC_ERR_INVALID, "k_correspondences must be >= 1");
  cudaStream_t st = c->stream;
  Scratch tmp(c);
  int* nbr = (int*)tmp.get(sizeof(int) * (size_t)k * cl.n);
  if (!cl.cov) cl.cov = (double*)c->get(sizeof(double) * 6 * (size_t)cl.n);
  if (!nbr || !cl.cov) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (covariances)");
  CK(c, cudaEventRecord(cl.ev[2], st));
  TRY(launch_knn_self(c, cl.view, cl.n, k, nbr, cl.d_tiles, cl.n_tiles, std::max(cl.n_clouds, 1)));
  CK(c, cudaEventRecord(cl.ev[3], st));
  c->mark("covariances: k-NN (tile + warp)");
  int min_cloud = cl.n;  // smallest cloud of a multi-cloud grid: every neighbour slot is filled iff it has >= k points
  for (int q = 0; q < cl.n_clouds; q++) min_cloud = std::min(min_cloud, cl.h_off[q + 1] - cl.h_off[q]);
  TRY(launch_covariance(c, st, cl.sorted, nbr, cl.n, cl.n, k, min_cloud >= k, method, cl.cov, nullptr, nullptr));  // full: the default k, every slot filled
  CK(c, cudaEventRecord(cl.ev[4], st));
  c->mark("covariances: k_covariance");
  cl.has_cov = true;
  cl.lazy_cov = false;
  cl.cov_speculative = speculative;
  cl.cov_k = k;
  cl.cov_method = method;
  cl.cov_timed = true;
  return RGC_OK;
}

// ------------------------------------------------------------------------------------------------
struct rgc_reg {
  rgc_ctx* ctx = nullptr;
  rgc_params prm{};
  Cloud src, tgt;
  // per-source-point state of the last linearize
  int* corr = nullptr;
  float* sqd = nullptr;
  double* maha = nullptr;
  double* partials = nullptr;      // per-block partial sums of k_linearize / k_compute_error (GICP path only)
  double* fit_partials = nullptr;  // k_fitness: 2 doubles per block (its own buffer: ADVICE r1, shared `partials` overflow)
  size_t fit_cap = 0;              // blocks
  int cap_src = 0;
  bool have_corr = false;
  // on-demand target covariances (FastGICP only): requests of the current linearize
  bool lazy_target = std::getenv("RGC_EAGER_TARGET_COV") == nullptr;
  int* need_list = nullptr;   // sorted target positions whose covariance is being computed
  int* need_count = nullptr;  // device counter (zeroed by k_linearize's last block)
  int* need_nbr = nullptr;    // k-major neighbour lists of need_list, stride cap_src
  int need_k = 0;
  // second set of per-point buffers + results of a linearize issued ahead of time (step_lm)
  int* corr2 = nullptr;
  float* sqd2 = nullptr;
  double* maha2 = nullptr;
  bool spec_ready = false;
  double spec_H[36], spec_b[6], spec_y0 = 0.0;
  int spec_inliers = 0;
  // LsqRegistration state
  double lm_lambda = -1.0;
  double final_hessian[36];  // row-major (symmetric)
  float final_T[16];         // row-major
  bool converged = false;
  int n_linearize = 0, n_compute_error = 0, last_inliers = 0;
  float lm_ms = 0;
  // voxelised GICP (FastVGICP)
  bool vgicp = false;
  double vox_res = 1.0;
  int vox_search = 2;  // DIRECT1
  int vox_mode = 0;    // ADDITIVE
  VoxelSlot* vox_slots = nullptr;
  VoxelMapView vox{};
  bool vox_valid = false;
  int vox_count = 0;
  int* vox_corr = nullptr;
  double* vox_maha = nullptr;
  double* vox_partials = nullptr;  // (148 * 8 + 8) x kLinN
  size_t vox_cap = 0;
  // sharded target (config C5)
  Slab slab{-1, 0.f, 0.f};
  rgc_comm* comm = nullptr;  // library-owned NCCL communicator: partial sums are all-reduced on the context's stream
  rgc_reduce_fn reduce_fn = nullptr;
  void* reduce_user = nullptr;
  double* reduce_buf = nullptr;  // device, caller-owned
};

// where the reduction kernels write, and (sharded) the cross-rank sum before the host reads it
constexpr int kSpecSlot = 32;  // doubles: results of the look-ahead linearize live at h_result + 32
static double* reg_result_ptr(rgc_reg* r) { return r->comm ? r->comm->d_buf : (r->reduce_fn ? r->reduce_buf : r->ctx->d_result); }
// where the look-ahead linearize (issued behind compute_error) puts its 29 doubles: right behind the error
// value in the communicator's buffer, so that ONE all-reduce sums both
static double* reg_spec_ptr(rgc_reg* r) { return r->comm ? r->comm->d_buf + 1 : r->ctx->d_result + kSpecSlot; }
// completion word for the next reduction kernel (null when the partial sums still have to cross ranks, or
// polling is disabled: the host then waits on the stream)
static DoneFlag reg_next_done(rgc_reg* r) {
  rgc_ctx* c = r->ctx;
  if (r->comm || r->reduce_fn || !c->spin_wait) return DoneFlag{nullptr, 0ull};
  return DoneFlag{c->d_seq, ++c->seq};
}
static int reg_spin(rgc_ctx* c) {
  // the last block of the last reduction kernel stored c->seq into mapped pinned memory after its
  // results: poll that word (an LM step is ~60 us of device work; a stream synchronize adds several us
  // of wake-up latency to every one of them)
  volatile unsigned long long* flag = c->h_seq;
  for (unsigned spins = 0; *flag != c->seq; spins++) {
    if ((spins & 0x3fff) == 0x3fff) {  // a failed launch / device fault never stores the word
      cudaError_t e = cudaStreamQuery(c->stream);
      if (e == cudaSuccess) break;
      if (e != cudaErrorNotReady) CK(c, e);
      cudaGetLastError();
    }
  }
  if (*flag != c->seq) CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}
// n_a doubles -> h_result[0..), n_b doubles (look-ahead linearize) -> h_result[kSpecSlot..)
static int reg_finish_reduce(rgc_reg* r, int n_a, int n_b = 0) {
  rgc_ctx* c = r->ctx;
  if (r->comm) {
    // one all-reduce of the n_a + n_b partial sums over NVLink that also moves the totals into the mapped host result
    // area and stores the completion word (k_peer_allreduce: one launch over peer memory; ncclAllReduce + k_publish
    // where the ranks' mailboxes could not be mapped): no host involvement in between
    const DoneFlag done = c->spin_wait ? DoneFlag{c->d_seq, ++c->seq} : DoneFlag{nullptr, 0ull};
    TRY(comm_allreduce(r->comm, n_a + n_b, true, n_a, c->d_result, n_b, c->d_result + kSpecSlot, done));
    if (c->spin_wait) return reg_spin(c);
  } else if (r->reduce_fn) {
    if (n_b) FAIL(c, RGC_ERR_STATE, "look-ahead is not available with an all-reduce callback");
    if (r->reduce_fn(r->reduce_user, r->reduce_buf, n_a) != 0) FAIL(c, RGC_ERR_STATE, "all-reduce hook failed");
    CK(c, cudaMemcpyAsync(c->h_result, r->reduce_buf, sizeof(double) * n_a, cudaMemcpyDeviceToHost, c->stream));
  } else if (c->spin_wait) {
    return reg_spin(c);
  }
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

static int reg_ensure_work(rgc_reg* r) {
  rgc_ctx* c = r->ctx;
  const int k = std::max(r->prm.k_correspondences, r->tgt.lazy_cov ? r->tgt.cov_k : 0);  // on-demand covariances keep their latched k
  if (r->cap_src >= r->src.n && r->corr && r->need_k >= k) return RGC_OK;
  c->put(r->corr);
  c->put(r->sqd);
  c->put(r->maha);
  c->put(r->corr2);
  c->put(r->sqd2);
  c->put(r->maha2);
  c->put(r->partials);
  c->put(r->need_list);
  c->put(r->need_nbr);
  r->cap_src = 0;
  const size_t n = (size_t)r->src.n;
  r->corr = (int*)c->get(4 * n);
  r->sqd = (float*)c->get(4 * n);
  r->maha = (double*)c->get(48 * n);
  r->corr2 = (int*)c->get(4 * n);
  r->sqd2 = (float*)c->get(4 * n);
  r->maha2 = (double*)c->get(48 * n);
  r->partials = (double*)c->get(sizeof(double) * kLinN * (size_t)reduce_grid(r->src.n));
  r->need_list = (int*)c->get(4 * n);
  r->need_nbr = (int*)c->get(4 * n * (size_t)k);
  if (!r->need_count) {
    r->need_count = (int*)c->get(4);
    if (r->need_count) CK(c, cudaMemsetAsync(r->need_count, 0, 4, c->stream));
  }
  if (!r->corr || !r->sqd || !r->maha || !r->corr2 || !r->sqd2 || !r->maha2 || !r->partials || !r->need_list || !r->need_nbr || !r->need_count)
    FAIL(c, RGC_ERR_NOMEM, "device allocation failed (work buffers)");
  r->cap_src = r->src.n;
  r->need_k = k;
  return RGC_OK;
}

// scratch of k_fitness only (it launches one block per 128 queries, unlike the persistent linearize grid)
static int reg_ensure_fitness(rgc_reg* r, int blocks) {
  rgc_ctx* c = r->ctx;
  if (r->fit_partials && r->fit_cap >= (size_t)blocks) return RGC_OK;
  c->put(r->fit_partials);
  r->fit_cap = 0;
  r->fit_partials = (double*)c->get(sizeof(double) * 2 * (size_t)blocks);
  if (!r->fit_partials) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (fitness scratch)");
  r->fit_cap = (size_t)blocks;
  return RGC_OK;
}

// can the target's covariances be computed on demand?  Only the exact-1-NN FastGICP path reads C_B at
// correspondences alone; the voxel map averages ALL of them, and covariances the user supplied or that
// were already computed for the whole cloud are simply used.
static bool target_lazy(const rgc_reg* r) { return r->lazy_target && !r->vgicp && !r->tgt.has_cov && r->prm.k_correspondences <= 32; }

static void to_rt(const double* T /*row-major 4x4*/, Rt& d, RtF& f) {
  for (int i = 0; i < 12; i++) {
    d.m[i] = T[i];
    f.m[i] = (float)T[i];  // Eigen::Isometry3d::cast<float>() (fast_gicp_impl.hpp:119)
  }
}

static int vgicp_n_off(const rgc_reg* r) { return r->vox_search == 2 ? 1 : (r->vox_search == 1 ? 7 : 27); }

// GaussianVoxelMap::create_voxelmap (fast_vgicp_voxel.hpp:129-156) on the device
static int vgicp_build(rgc_reg* r) {
  rgc_ctx* c = r->ctx;
  if (r->vox_valid) return RGC_OK;
  Cloud& t = r->tgt;
  const int n = t.n;
  cudaStream_t st = c->stream;
  VoxGeom g;
  g.res = r->vox_res;
  int total_bits = 0;
  for (int a = 0; a < 3; a++) {
    const int lo = (int)std::floor((double)t.bb_min[a] / g.res - 0.5), hi = (int)std::floor((double)t.bb_max[a] / g.res - 0.5);
    g.lo[a] = lo;
    g.dim[a] = hi - lo + 1;
    int b = 1;
    while ((1 << b) < g.dim[a]) b++;
    g.bits[a] = b;
    total_bits += b;
  }
  if (total_bits > 62) FAIL(c, RGC_ERR_UNSUPPORTED, "voxel grid too large for a 62-bit key (resolution too small for this extent)");
  const size_t n_sz = (size_t)n;
  Scratch tmp(c);
  uint64_t* keys_a = (uint64_t*)tmp.get(8 * n_sz);
  uint64_t* keys_b = (uint64_t*)tmp.get(8 * n_sz);
  uint32_t* vals_a = (uint32_t*)tmp.get(4 * n_sz);
  uint32_t* vals_b = (uint32_t*)tmp.get(4 * n_sz);
  uint32_t* hist = (uint32_t*)tmp.get(4 * rs_scratch_words(n, RS_MAX_PASSES));
  unsigned int* d_cnt = (unsigned int*)tmp.get(4);
  int* heads = (int*)tmp.get(4 * n_sz);
  if (!keys_a || !keys_b || !vals_a || !vals_b || !hist || !d_cnt || !heads) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (voxel map)");
  k_vox_keys<<<div_up(n, 256), 256, 0, st>>>(t.sorted, n, g, keys_a, vals_a);
  CKL(c);
  uint64_t* ks = keys_a;
  uint32_t* vs = vals_a;
  TRY(radix_sort_pairs(c, keys_a, keys_b, vals_a, vals_b, hist, n, total_bits, &ks, &vs));
  CK(c, cudaMemsetAsync(d_cnt, 0, 4, st));
  k_vox_count<<<div_up(n, 256), 256, 0, st>>>(ks, n, d_cnt, heads);
  CKL(c);
  CK(c, cudaMemcpyAsync(c->h_counts, d_cnt, 4, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  const int nv = (int)c->h_counts[0];
  size_t slots = 8;
  int lg = 3;
  while (slots < 2 * (size_t)nv) {
    slots <<= 1;
    lg++;
  }
  c->put(r->vox_slots);
  r->vox_slots = (VoxelSlot*)c->get(slots * sizeof(VoxelSlot));
  if (!r->vox_slots) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (voxel table)");
  CK(c, cudaMemsetAsync(r->vox_slots, 0xff, slots * sizeof(VoxelSlot), st));
  k_vox_reduce<<<div_up(nv * 32, 128), 128, 0, st>>>(ks, vs, n, heads, nv, t.sorted, t.inv, t.cov, r->vox_mode, r->vox_slots, (uint32_t)(slots - 1),
                                                     (uint32_t)(64 - lg));
  CKL(c);
  r->vox = VoxelMapView{r->vox_slots, (uint32_t)(slots - 1), (uint32_t)(64 - lg), g};
  r->vox_count = nv;
  r->vox_valid = true;
  return RGC_OK;
}

static int vgicp_ensure_work(rgc_reg* r) {
  rgc_ctx* c = r->ctx;
  const size_t need = (size_t)r->src.n * vgicp_n_off(r);
  if (r->vox_cap >= need && r->vox_corr) return RGC_OK;
  c->put(r->vox_corr);
  c->put(r->vox_maha);
  r->vox_cap = 0;
  r->vox_corr = (int*)c->get(4 * need);
  r->vox_maha = (double*)c->get(48 * need);
  if (!r->vox_partials) r->vox_partials = (double*)c->get(sizeof(double) * kLinN * (size_t)(148 * 8 + 8));
  if (!r->vox_corr || !r->vox_maha || !r->vox_partials) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (vgicp work buffers)");
  r->vox_cap = need;
  return RGC_OK;
}

// FastVGICP::linearize (fast_vgicp_impl.hpp:119-180)
static int vgicp_linearize(rgc_reg* r, const double* T, double* err, double* H, double* b) {
  rgc_ctx* c = r->ctx;
  TRY(vgicp_build(r));
  TRY(vgicp_ensure_work(r));
  Rt Td;
  RtF Tf;
  to_rt(T, Td, Tf);
  const int want = (H && b) ? 1 : 0;
  const int n_off = vgicp_n_off(r);
  const long long total = (long long)r->src.n * n_off;
  const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 8);
  k_vgicp_linearize<<<grid, kThreads, 0, c->stream>>>(r->vox, r->src.sorted, r->src.cov, r->src.n, r->vox_search, n_off, Td, want, r->vox_corr, r->vox_maha,
                                                     r->vox_partials, c->d_ticket, reg_result_ptr(r), reg_next_done(r));
  CKL(c);
  TRY(reg_finish_reduce(r, kLinN));
  r->n_linearize++;
  r->have_corr = true;
  const double* res = c->h_result;
  *err = res[0];
  r->last_inliers = (int)res[kAccN];
  if (want) {
    int o = 1;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) {
        H[i * 6 + j] = H[j * 6 + i] = res[o];
        o++;
      }
    for (int i = 0; i < 6; i++) b[i] = res[22 + i];
  }
  return RGC_OK;
}

// FastVGICP::compute_error (fast_vgicp_impl.hpp:183-204)
static int vgicp_compute_error(rgc_reg* r, const double* T, double* err) {
  rgc_ctx* c = r->ctx;
  if (!r->have_corr || !r->vox_valid) FAIL(c, RGC_ERR_STATE, "compute_error before any linearize");
  Rt Td;
  RtF Tf;
  to_rt(T, Td, Tf);
  const int n_off = vgicp_n_off(r);
  const long long total = (long long)r->src.n * n_off;
  const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 8);
  k_vgicp_compute_error<<<grid, kThreads, 0, c->stream>>>(r->vox, r->src.sorted, r->src.n, n_off, Td, r->vox_corr, r->vox_maha, r->vox_partials, c->d_ticket,
                                                         reg_result_ptr(r), reg_next_done(r));
  CKL(c);
  TRY(reg_finish_reduce(r, 1));
  r->n_compute_error++;
  *err = c->h_result[0];
  return RGC_OK;
}

// FastGICP::linearize (fast_gicp_impl.hpp:155-211), launch half: correspondences (seeded by `hint`)
// into corr/sqd, then the per-point terms and their reduction into `result`
struct CeJob {  // compute_error at the same pose, fused into the correspondence launch (k_trial_step)
  bool on = false;
  double* result = nullptr;
  DoneFlag done{nullptr, 0ull};
};
static int gicp_linearize_launch(rgc_reg* r, const double* T, int want, const int* hint, int* corr, float* sqd, double* maha, double* result,
                                 const CeJob& ce = CeJob()) {
  rgc_ctx* c = r->ctx;
  Rt Td;
  RtF Tf;
  to_rt(T, Td, Tf);
  const float thr = r->prm.max_correspondence_distance;
  const float thr2 = thr * thr;  // float product, +inf for the FLT_MAX default (fast_gicp_impl.hpp:136)
  const int spread = query_spread(r->src.n);
  const bool lazy = r->tgt.lazy_cov;
  if (c->profile) CK(c, cudaEventRecord(c->evk[0], c->stream));
  // (a warp-cooperative variant of this search, one warp per 32 Morton-adjacent source points, was exact
  // too but slower on one sweep, 155 vs 92 us: profiles/README.md; it was removed)
  const int corr_blocks = div_up(r->src.n * spread, kThreads);
  if (ce.on) {
    const int ce_blocks = reduce_grid(r->src.n);
    launch_pdl(c, k_trial_step, dim3(ce_blocks + corr_blocks), dim3(kThreads), 0, c->stream, r->tgt.view, r->src.sorted, r->src.n, spread, Tf, thr2, r->slab, hint, corr, sqd,
                                                                      lazy ? r->tgt.cov_state : nullptr, r->need_list, r->need_count, ce_blocks, Td, r->maha,
                                                                      r->partials, c->d_ticket, ce.result, ce.done);
  } else {
    launch_pdl(c, k_correspond, dim3(corr_blocks), dim3(kThreads), 0, c->stream, r->tgt.view, r->src.sorted, r->src.n, spread, Tf, thr2, r->slab, hint, corr, sqd,
                                                          lazy ? r->tgt.cov_state : nullptr, r->need_list, r->need_count);
  }
  CKL(c);
  c->mark(ce.on ? "lm: k_trial_step (compute_error + correspond)" : "lm: k_correspond");
  if (c->profile) CK(c, cudaEventRecord(c->evk[1], c->stream));
  if (lazy) {
    // on-demand target covariances: exact k-NN (one warp per query) + covariance of exactly the target
    // points that became correspondences for the first time.  The launches are sized for the worst case
    // (every source point a new target point) and read the real count from the device: nothing here
    // makes the host wait.  After the first linearize of an align the list is nearly empty.
    const int k = r->tgt.cov_k, method = r->tgt.cov_method, cap = r->cap_src, n_t = r->tgt.n;
    launch_pdl(c, k_knn_warp, dim3(std::min(div_up(r->src.n, KW_WARPS), 148 * RGC_KW_MINB)), dim3(KW_WARPS * 32), 0, c->stream, r->tgt.view, n_t, k, r->need_count, nullptr, r->need_list, cap, r->need_nbr, nullptr, nullptr, 0);
    CKL(c);
    c->mark("lm: on-demand k_knn_warp");
    TRY(launch_covariance(c, c->stream, r->tgt.sorted, r->need_nbr, cap, r->src.n, k, n_t >= k, method, r->tgt.cov, r->need_list, r->need_count));
    c->mark("lm: on-demand k_covariance");
    if (c->profile) CK(c, cudaEventRecord(c->evk[3], c->stream));
  }
  TRY(join_side(c));  // the source covariances (lane 1) are first read here: reg_ready joined the source's build only
  launch_pdl(c, k_linearize, dim3(reduce_grid(r->src.n)), dim3(kThreads), 0, c->stream, r->tgt.sorted, r->src.sorted, r->src.cov, r->tgt.cov, r->src.n, Td, want, corr, maha, r->partials,
                                                                     c->d_ticket, result, reg_next_done(r), lazy ? r->need_count : nullptr);
  CKL(c);
  c->mark("lm: k_linearize");
  if (c->profile) CK(c, cudaEventRecord(c->evk[2], c->stream));
  return RGC_OK;
}
static void gicp_linearize_unpack(const double* res, double* err, int* inliers, double* H, double* b) {
  *err = res[0];
  *inliers = (int)res[kAccN];
  if (H && b) {
    int o = 1;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) {
        H[i * 6 + j] = H[j * 6 + i] = res[o];
        o++;
      }
    for (int i = 0; i < 6; i++) b[i] = res[22 + i];
  }
}

// H row-major 6x6
static int reg_linearize(rgc_reg* r, const double* T, double* err, double* H, double* b) {
  rgc_ctx* c = r->ctx;
  if (r->vgicp) return vgicp_linearize(r, T, err, H, b);
  TRY(reg_ensure_work(r));
  TRY(gicp_linearize_launch(r, T, (H && b) ? 1 : 0, r->have_corr ? r->corr : nullptr, r->corr, r->sqd, r->maha, reg_result_ptr(r)));
  TRY(reg_finish_reduce(r, kLinN));
  if (c->profile) {
    cudaEventSynchronize(c->evk[2]);
    cudaEventElapsedTime(&c->last_kernel_ms[0], c->evk[0], c->evk[1]);
    if (r->tgt.lazy_cov) {
      cudaEventElapsedTime(&c->last_ondemand_ms, c->evk[1], c->evk[3]);
      cudaEventElapsedTime(&c->last_kernel_ms[1], c->evk[3], c->evk[2]);
    } else {
      c->last_ondemand_ms = 0.f;
      cudaEventElapsedTime(&c->last_kernel_ms[1], c->evk[1], c->evk[2]);
    }
  }
  r->n_linearize++;
  r->have_corr = true;
  gicp_linearize_unpack(c->h_result, err, &r->last_inliers, H, b);
  return RGC_OK;
}

// FastGICP::compute_error (fast_gicp_impl.hpp:214-237)
// With `ahead`, the linearize at the same pose is issued right behind the error kernel, into the
// second buffer set, and both are collected with ONE wait: if the LM step is then accepted (the
// usual case) the next iteration's H, b are already on the host.  lsq_registration_impl.hpp:125-172
// calls compute_error(xi) and, after an accepted step, linearize(xi) at the start of the next
// iteration: same kernels, same inputs, same results, one host round trip less per iteration.
static int reg_compute_error(rgc_reg* r, const double* T, double* err, bool ahead = false) {
  rgc_ctx* c = r->ctx;
  if (r->vgicp) return vgicp_compute_error(r, T, err);
  if (!r->have_corr) FAIL(c, RGC_ERR_STATE, "compute_error before any linearize");
  Rt Td;
  RtF Tf;
  to_rt(T, Td, Tf);
  if (c->profile) CK(c, cudaEventRecord(c->evk[0], c->stream));
  if (ahead && c->fuse_trial) {
    // the error blocks ride in the correspondence launch of the look-ahead linearize (k_trial_step)
    CeJob ce;
    ce.on = true;
    ce.result = reg_result_ptr(r);
    ce.done = reg_next_done(r);
    TRY(gicp_linearize_launch(r, T, 1, r->corr, r->corr2, r->sqd2, r->maha2, reg_spec_ptr(r), ce));
  } else {
    launch_pdl(c, k_compute_error, dim3(reduce_grid(r->src.n)), dim3(kThreads), 0, c->stream, r->tgt.sorted, r->src.sorted, r->src.n, Td, r->corr, r->maha, r->partials,
                                                                           c->d_ticket, reg_result_ptr(r), reg_next_done(r));
    CKL(c);
    if (c->profile) CK(c, cudaEventRecord(c->evk[1], c->stream));
    if (ahead) TRY(gicp_linearize_launch(r, T, 1, r->corr, r->corr2, r->sqd2, r->maha2, reg_spec_ptr(r)));
  }
  TRY(reg_finish_reduce(r, 1, ahead ? kLinN : 0));
  if (c->profile) {
    cudaEventSynchronize(c->evk[1]);
    cudaEventElapsedTime(&c->last_kernel_ms[2], c->evk[0], c->evk[1]);
  }
  r->n_compute_error++;
  *err = c->h_result[0];
  if (ahead) gicp_linearize_unpack(c->h_result + kSpecSlot, &r->spec_y0, &r->spec_inliers, r->spec_H, r->spec_b);
  return RGC_OK;
}
// an accepted step makes the look-ahead linearize the current one
static void reg_adopt_ahead(rgc_reg* r) {
  std::swap(r->corr, r->corr2);
  std::swap(r->sqd, r->sqd2);
  std::swap(r->maha, r->maha2);
  r->last_inliers = r->spec_inliers;
  r->n_linearize++;
  r->spec_ready = true;
}

// what follows a finished build: the covariances are STARTED right away (fast_gicp_impl.hpp:104-109
// computes them at the first align; they are redone if k / regularisation change before that), except a
// target whose covariances are computed on demand; a source build ends with the event the main stream joins
static int post_build(rgc_reg* r, Cloud& cl) {
  rgc_ctx* c = r->ctx;
  int rc = RGC_OK;
  if (!(&cl == &r->tgt && target_lazy(r))) rc = cloud_covariances(c, cl, r->prm.k_correspondences, r->prm.regularization, true);
  if (c->lane == 1) {
    cudaEventRecord(c->join_ev, c->stream);
    c->side_pending = true;
  }
  return rc;
}

// run the next phase of a cloud's build (its host wait has been satisfied or is about to be)
static int build_advance(rgc_reg* r, Cloud& cl) {
  rgc_ctx* c = r->ctx;
  LaneScope ls(c, cl.job->lane);
  int rc = cl.job->stage == 1 ? build_phase2(c, cl) : (cl.job->stage == 3 ? build_passes(c, cl) : build_phase3(c, cl));
  if (rc == RGC_OK && !cl.job) rc = post_build(r, cl);
  if (rc != RGC_OK) cloud_release(c, cl);
  return rc;
}

// complete the builds in flight.  Whichever cloud's awaited host copy has landed goes first, so while the
// host sleeps on one cloud's copy the device works on the other cloud's kernels.
static int reg_drain(rgc_reg* r) {
  rgc_ctx* c = r->ctx;
  // sort passes held back by the setters: the source's first (build_phase1)
  for (Cloud* cl : {&r->src, &r->tgt})
    if (cl->job && cl->job->stage == 3) TRY(build_advance(r, *cl));
  for (;;) {
    Cloud* cls[2] = {&r->tgt, &r->src};
    bool pending = false, progressed = false;
    for (Cloud* cl : cls) {
      if (!cl->job) continue;
      pending = true;
      const cudaError_t q = cudaEventQuery(cl->job->ready);
      if (q == cudaSuccess) {
        TRY(build_advance(r, *cl));
        progressed = true;
      } else if (q == cudaErrorNotReady) {
        cudaGetLastError();  // not an error
      } else {
        CK(c, q);
      }
    }
    if (!pending) return RGC_OK;
    if (!progressed) {
      Cloud* w = r->tgt.job ? &r->tgt : &r->src;
      TRY(build_advance(r, *w));  // blocks on that cloud's event
    }
  }
}

static int reg_ready(rgc_reg* r) {
  rgc_ctx* c = r->ctx;
  TRY(reg_drain(r));
  if (!r->src.valid || !r->tgt.valid) FAIL(c, RGC_ERR_STATE, "source and target clouds must both be set");
  // FastGICP with the source prepared on lane 1 and its covariances already under way with the right
  // parameters: the main stream only waits for the source's BUILD, so the first correspondence search and the
  // on-demand target covariances overlap the source's k-NN; k_linearize joins the rest (gicp_linearize_launch)
  {
    const Cloud& s = r->src;
    const bool cov_ok = s.has_cov && !(s.cov_speculative && (s.cov_k != r->prm.k_correspondences || s.cov_method != r->prm.regularization));
    if (c->late_join && !r->vgicp && cov_ok && s.lane_built == 1 && r->tgt.lane_built == 0)
      TRY(join_side_build(c, s));
    else
      TRY(join_side(c));
  }
  // fast_gicp_impl.hpp:104-109 — covariances are computed lazily, source first (normally both
  // were already started by set_input, see set_cloud)
  TRY(cloud_covariances(c, r->src, r->prm.k_correspondences, r->prm.regularization));
  if (target_lazy(r)) {
    // Target covariances on demand: linearize reads target_covs_[target_index] only at the current
    // correspondences (fast_gicp_impl.hpp:139-146), <= n_source of the n_target covariances the
    // reference computes up front (:107-109).  Each one is still the covariance of that point's exact
    // k nearest neighbours with the parameters in force at the first align, computed by the same
    // kernels: the values linearize sees are bit-identical to the eager pass.
    Cloud& t = r->tgt;
    if (!t.lazy_cov) {
      const int k = r->prm.k_correspondences;
      if (k < 1) FAIL(c, RGC_ERR_INVALID, "k_correspondences must be >= 1");
      if (!t.cov) t.cov = (double*)c->get(sizeof(double) * 6 * (size_t)t.n);
      if (!t.cov_state) t.cov_state = (int*)c->get(sizeof(int) * (size_t)t.n);
      if (!t.cov || !t.cov_state) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (on-demand covariances)");
      CK(c, cudaMemsetAsync(t.cov_state, 0, sizeof(int) * (size_t)t.n, c->stream));
      t.cov_k = k;  // latched like the reference's covariances: computed once per cloud
      t.cov_method = r->prm.regularization;
      t.lazy_cov = true;
    }
  } else {
    TRY(cloud_covariances(c, r->tgt, r->prm.k_correspondences, r->prm.regularization));
  }
  return RGC_OK;
}

// lsq_registration_impl.hpp:106-122
static int step_gn(rgc_reg* r, double* x0, double* delta) {
  double H[36], b[6], nb[6], d[6], y0;
  TRY(reg_linearize(r, x0, &y0, H, b));
  for (int i = 0; i < 6; i++) nb[i] = -b[i];
  lm::solve_ldlt6(H, nb, d);
  lm::se3_delta(d, delta);
  lm::mul4(delta, x0, x0);
  std::memcpy(r->final_hessian, H, sizeof(H));
  return 1;
}

// lsq_registration_impl.hpp:125-172 ; returns 1 = step taken, 0 = "lm not converged", <0 = error
static int step_lm(rgc_reg* r, double* x0, double* delta, double* y0_out, bool more_iterations) {
  double H[36], b[6], y0;
  if (r->spec_ready) {  // linearized at x0 behind the previous step's compute_error
    std::memcpy(H, r->spec_H, sizeof(H));
    std::memcpy(b, r->spec_b, sizeof(b));
    y0 = r->spec_y0;
    r->spec_ready = false;
  } else {
    TRY(reg_linearize(r, x0, &y0, H, b));
  }
  *y0_out = y0;
  if (r->lm_lambda < 0.0) {
    double mx = 0.0;
    for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i * 7]));
    r->lm_lambda = r->prm.lm_init_lambda_factor * mx;
  }
  // the look-ahead needs the plain single-GPU GICP path (no all-reduce hook, no per-kernel timing)
  const bool can_look_ahead = more_iterations && r->ctx->look_ahead && !r->vgicp && !r->reduce_fn && !r->ctx->profile;
  double nu = 2.0;
  for (int i = 0; i < r->prm.lm_max_iterations; i++) {
    double A[36], nb[6], d[6], xi[16], yi;
    for (int j = 0; j < 36; j++) A[j] = H[j];
    for (int j = 0; j < 6; j++) {
      A[j * 7] += r->lm_lambda;
      nb[j] = -b[j];
    }
    lm::solve_ldlt6(A, nb, d);
    lm::se3_delta(d, delta);
    lm::mul4(delta, x0, xi);
    // if this step is accepted and does not end the outer loop, the next thing needed is linearize(xi)
    const bool ahead = can_look_ahead && !lm::is_converged(delta, r->prm.rotation_epsilon, r->prm.transformation_epsilon);
    TRY(reg_compute_error(r, xi, &yi, ahead));
    double denom = 0.0;
    for (int j = 0; j < 6; j++) denom += d[j] * (r->lm_lambda * d[j] - b[j]);
    const double rho = (y0 - yi) / denom;
    if (r->prm.lm_debug_print) {
      if (i == 0) std::printf("--- LM optimization ---\n%5s %15s %15s %15s %15s %15s %5s\n", "i", "y0", "yi", "rho", "lambda", "|delta|", "dec");
      double dn = 0.0;
      for (int j = 0; j < 6; j++) dn += d[j] * d[j];
      std::printf("%5d %15g %15g %15g %15g %15g %5c\n", i, y0, yi, rho, r->lm_lambda, std::sqrt(dn), rho > 0.0 ? 'x' : ' ');
    }
    if (rho < 0) {
      if (lm::is_converged(delta, r->prm.rotation_epsilon, r->prm.transformation_epsilon)) return 1;
      r->lm_lambda = nu * r->lm_lambda;
      nu = 2 * nu;
      continue;
    }
    std::memcpy(x0, xi, sizeof(xi));
    r->lm_lambda = r->lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    std::memcpy(r->final_hessian, H, sizeof(H));
    if (ahead) reg_adopt_ahead(r);
    return 1;
  }
  return 0;
}

// ================================================================================================
extern "C" {

int rgc_ctx_create(int device, rgc_ctx** out) {
  if (!out) return RGC_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    cudaGetLastError();
    std::fprintf(stderr, "rgc_ctx_create: no usable CUDA device %d (found %d); this library has no CPU fallback\n", device, count);
    return RGC_ERR_CUDA;
  }
  rgc_ctx* c = new rgc_ctx();
  c->device = device;
  bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaHostAlloc((void**)&c->h_result, sizeof(double) * 64, cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer((void**)&c->d_result, c->h_result, 0) == cudaSuccess &&
            cudaHostAlloc((void**)&c->h_bbox, sizeof(float) * 6 * kBboxBlocks, cudaHostAllocDefault) == cudaSuccess &&
            cudaHostAlloc((void**)&c->h_counts, sizeof(uint32_t) * kMaxLevels, cudaHostAllocDefault) == cudaSuccess &&
            cudaMalloc((void**)&c->d_ticket, 64) == cudaSuccess && cudaMemset(c->d_ticket, 0, 64) == cudaSuccess &&
            cudaHostAlloc((void**)&c->h_seq, 64, cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer((void**)&c->d_seq, c->h_seq, 0) == cudaSuccess;
  if (ok) *c->h_seq = 0ull;
  // lane 1 (source cloud: ~25 small kernels) gets the higher priority, so that its blocks are
  // dispatched ahead of the thousands of pending blocks of the target's kNN launch on lane 0
  // (without it the source build took 0.87 ms instead of 0.19 ms when overlapped)
  int prio_lo = 0, prio_hi = 0;
  ok = ok && cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) == cudaSuccess &&
       cudaStreamCreateWithPriority(&c->parked.stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
       cudaHostAlloc((void**)&c->parked.h_bbox, sizeof(float) * 6 * kBboxBlocks, cudaHostAllocDefault) == cudaSuccess &&
       cudaHostAlloc((void**)&c->parked.h_counts, sizeof(uint32_t) * kMaxLevels, cudaHostAllocDefault) == cudaSuccess &&
       cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming) == cudaSuccess &&
       cudaStreamCreateWithPriority(&c->aux[0], cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
       cudaStreamCreateWithPriority(&c->aux[1], cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  for (int i = 0; ok && i < 8; i++) ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  for (int i = 0; ok && i < 4; i++) ok = cudaEventCreate(&c->evk[i]) == cudaSuccess;
  c->profile = std::getenv("RGC_PROFILE") != nullptr;
  if (!ok) {
    std::fprintf(stderr, "rgc_ctx_create: %s\n", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return RGC_ERR_CUDA;
  }
  *out = c;
  return RGC_OK;
}

int rgc_ctx_destroy(rgc_ctx* c) {
  if (!c) return RGC_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->parked.stream);
  for (cudaStream_t a : c->aux)
    if (a) {
      cudaStreamSynchronize(a);
      cudaStreamDestroy(a);
    }
  for (auto& kv : c->block_info) cudaFree(kv.first);
  cudaFreeHost(c->h_result);
  cudaFreeHost(c->h_seq);
  cudaFreeHost(c->h_bbox);
  cudaFreeHost(c->h_counts);
  cudaFreeHost(c->parked.h_bbox);
  cudaFreeHost(c->parked.h_counts);
  cudaEventDestroy(c->join_ev);
  for (cudaEvent_t e : c->free_events) cudaEventDestroy(e);
  for (float* p : c->free_hslots) cudaFreeHost(p);
  cudaStreamDestroy(c->parked.stream);
  cudaFree(c->d_ticket);
  for (int i = 0; i < 8; i++) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 4; i++) cudaEventDestroy(c->evk[i]);
  cudaStreamDestroy(c->stream);
  delete c;
  return RGC_OK;
}

const char* rgc_last_error(const rgc_ctx* c) { return c ? c->err.c_str() : "null context"; }
int rgc_ctx_synchronize(rgc_ctx* c) {
  CK(c, cudaStreamSynchronize(c->parked.stream));
  CK(c, cudaStreamSynchronize(c->stream));
  for (cudaStream_t a : c->aux) CK(c, cudaStreamSynchronize(a));
  return RGC_OK;
}
void* rgc_ctx_stream(rgc_ctx* c) { return (void*)c->stream; }
int rgc_ctx_set_profiling(rgc_ctx* c, int on) {
  if (!c) return RGC_ERR_INVALID;
  c->profile = on != 0;
  return RGC_OK;
}
int rgc_ctx_last_kernel_ms(const rgc_ctx* c, float* ms3) {
  if (!c || !ms3) return RGC_ERR_INVALID;
  for (int i = 0; i < 3; i++) ms3[i] = c->last_kernel_ms[i];
  return RGC_OK;
}
uint64_t rgc_ctx_launch_count(const rgc_ctx* c) { return c->launches; }

void rgc_params_default(rgc_params* p) {
  p->max_iterations = 64;
  p->rotation_epsilon = 2e-3;
  p->transformation_epsilon = 5e-4;
  p->max_correspondence_distance = FLT_MAX;
  p->k_correspondences = 20;
  p->regularization = RGC_REG_PLANE;
  p->optimizer = RGC_OPT_LEVENBERG_MARQUARDT;
  p->lm_max_iterations = 10;
  p->lm_init_lambda_factor = 1e-9;
  p->lm_debug_print = 0;
  p->grid_cell = 0.f;
}

int rgc_reg_create(rgc_ctx* c, rgc_reg** out) {
  if (!c || !out) return RGC_ERR_INVALID;
  rgc_reg* r = new rgc_reg();
  r->ctx = c;
  rgc_params_default(&r->prm);
  for (int i = 0; i < 36; i++) r->final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;  // final_hessian_.setIdentity()
  for (int i = 0; i < 16; i++) r->final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  *out = r;
  return RGC_OK;
}

int rgc_reg_destroy(rgc_reg* r) {
  if (!r) return RGC_OK;
  rgc_ctx* c = r->ctx;
  cudaSetDevice(c->device);
  cloud_release(c, r->src);
  cloud_release(c, r->tgt);
  c->put(r->corr);
  c->put(r->sqd);
  c->put(r->maha);
  c->put(r->corr2);
  c->put(r->sqd2);
  c->put(r->maha2);
  c->put(r->partials);
  c->put(r->fit_partials);
  c->put(r->need_list);
  c->put(r->need_count);
  c->put(r->need_nbr);
  c->put(r->vox_slots);
  c->put(r->vox_corr);
  c->put(r->vox_maha);
  c->put(r->vox_partials);
  delete r;
  return RGC_OK;
}

int rgc_reg_set_params(rgc_reg* r, const rgc_params* p) {
  if (!r || !p) return RGC_ERR_INVALID;
  if (p->k_correspondences < 1 || p->k_correspondences > kMaxK) FAIL(r->ctx, RGC_ERR_UNSUPPORTED, "k_correspondences must be in [1, 128]");
  if (p->regularization < 0 || p->regularization > 4) FAIL(r->ctx, RGC_ERR_INVALID, "unknown regularization method");
  r->prm = *p;
  return RGC_OK;
}
int rgc_reg_get_params(const rgc_reg* r, rgc_params* p) {
  if (!r || !p) return RGC_ERR_INVALID;
  *p = r->prm;
  return RGC_OK;
}

static int set_cloud(rgc_reg* r, Cloud& cl, const void* pts, size_t n, size_t stride, uint64_t key, bool on_device) {
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  if (key != 0 && (cl.valid || cl.job) && (cl.job ? cl.job->key : cl.key) == key) return RGC_OK;  // fast_gicp_impl.hpp:73-75 / :84-86
  r->have_corr = false;
  if (&cl == &r->tgt) r->vox_valid = false;  // FastVGICP::setInputTarget resets the voxel map (fast_vgicp_impl.hpp:57-64)
  // The source cloud is prepared on lane 1, the target on the main stream.  Only the first phase of the
  // build (upload, ingest) is issued here; the rest, and the covariances, follow in reg_drain.
  {
    LaneScope ls(c, (&cl == &r->src && c->overlap) ? 1 : 0);
    // (holding the target's sort passes back until the drain so that the source's are issued first was measured:
    // the target build then starts ~60 us later and the cold align got 25 us slower; the passes go out right away)
    int rc = build_phase1(c, cl, pts, n, stride, on_device, key, r->prm.grid_cell, nullptr, 0, false);
    if (rc != RGC_OK) {
      cloud_release(c, cl);
      return rc;
    }
  }
  return c->defer_builds ? RGC_OK : reg_drain(r);
}
int rgc_reg_set_source(rgc_reg* r, const void* p, size_t n, size_t s, uint64_t key) { return r ? set_cloud(r, r->src, p, n, s, key, false) : RGC_ERR_INVALID; }
int rgc_reg_set_target(rgc_reg* r, const void* p, size_t n, size_t s, uint64_t key) { return r ? set_cloud(r, r->tgt, p, n, s, key, false) : RGC_ERR_INVALID; }
int rgc_reg_set_source_device(rgc_reg* r, const void* p, size_t n, size_t s, uint64_t key) { return r ? set_cloud(r, r->src, p, n, s, key, true) : RGC_ERR_INVALID; }
int rgc_reg_set_target_device(rgc_reg* r, const void* p, size_t n, size_t s, uint64_t key) { return r ? set_cloud(r, r->tgt, p, n, s, key, true) : RGC_ERR_INVALID; }

int rgc_reg_sync_inputs(rgc_reg* r) {
  if (!r) return RGC_ERR_INVALID;
  CK(r->ctx, cudaSetDevice(r->ctx->device));
  return reg_drain(r);
}
int rgc_reg_swap_source_and_target(rgc_reg* r) {
  if (!r) return RGC_ERR_INVALID;
  TRY(reg_drain(r));
  std::swap(r->src, r->tgt);
  r->have_corr = false;  // correspondences_.clear(); sq_distances_.clear();
  r->vox_valid = false;  // voxelmap_.reset() (fast_vgicp_impl.hpp:46-54)
  return RGC_OK;
}
int rgc_reg_clear_source(rgc_reg* r) {
  if (!r) return RGC_ERR_INVALID;
  cloud_release(r->ctx, r->src);
  r->have_corr = false;
  return RGC_OK;
}
int rgc_reg_clear_target(rgc_reg* r) {
  if (!r) return RGC_ERR_INVALID;
  cloud_release(r->ctx, r->tgt);
  r->have_corr = false;
  r->vox_valid = false;
  return RGC_OK;
}

static int set_covs(rgc_reg* r, Cloud& cl, const double* m, size_t n) {
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  if (!cl.valid) FAIL(c, RGC_ERR_STATE, "set the point cloud before its covariances");
  if ((int)n != cl.n) FAIL(c, RGC_ERR_INVALID, "covariance count does not match the cloud size");
  TRY(join_side(c));
  Scratch tmp(c);
  double* stage = (double*)tmp.get(128 * n);
  if (!cl.cov) cl.cov = (double*)c->get(48 * n);
  if (!stage || !cl.cov) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (covariance import)");
  CK(c, cudaMemcpyAsync(stage, m, 128 * n, cudaMemcpyHostToDevice, c->stream));
  k_cov_import<<<div_up(cl.n, 256), 256, 0, c->stream>>>(cl.sorted, cl.n, stage, cl.cov);
  CKL(c);
  CK(c, cudaStreamSynchronize(c->stream));
  cl.has_cov = true;
  cl.lazy_cov = false;
  cl.cov_speculative = false;  // user-provided: kept whatever the parameters
  cl.cov_timed = false;
  cl.knn_ms = cl.cov_ms = 0.f;
  if (&cl == &r->tgt) r->vox_valid = false;
  return RGC_OK;
}
static int get_covs(rgc_reg* r, Cloud& cl, double* m, size_t n) {
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  if (!cl.valid) FAIL(c, RGC_ERR_STATE, "no point cloud set");
  if ((int)n != cl.n) FAIL(c, RGC_ERR_INVALID, "covariance count does not match the cloud size");
  TRY(join_side(c));
  TRY(cloud_covariances(c, cl, r->prm.k_correspondences, r->prm.regularization));
  Scratch tmp(c);
  double* stage = (double*)tmp.get(128 * n);
  if (!stage) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (covariance export)");
  k_cov_export<<<div_up(cl.n, 256), 256, 0, c->stream>>>(cl.sorted, cl.n, cl.cov, stage);
  CKL(c);
  CK(c, cudaMemcpyAsync(m, stage, 128 * n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}
int rgc_reg_set_source_covs(rgc_reg* r, const double* m, size_t n) { return (r && m) ? set_covs(r, r->src, m, n) : RGC_ERR_INVALID; }
int rgc_reg_set_target_covs(rgc_reg* r, const double* m, size_t n) { return (r && m) ? set_covs(r, r->tgt, m, n) : RGC_ERR_INVALID; }
int rgc_reg_get_source_covs(rgc_reg* r, double* m, size_t n) { return (r && m) ? get_covs(r, r->src, m, n) : RGC_ERR_INVALID; }
int rgc_reg_get_target_covs(rgc_reg* r, double* m, size_t n) { return (r && m) ? get_covs(r, r->tgt, m, n) : RGC_ERR_INVALID; }

static void colmajor_to_row(const double* in, double* out) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[r * 4 + c] = in[c * 4 + r];
}

int rgc_reg_align(rgc_reg* r, const float* guess, float* final_T16, rgc_result* result, float* out_points) {
  if (!r) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  CK(c, cudaEventRecord(c->ev[5], c->stream));
  c->mark("align: begin");
  TRY(reg_ready(r));
  // lsq_registration_impl.hpp:53-57
  double x0[16];
  for (int rr = 0; rr < 4; rr++)
    for (int cc = 0; cc < 4; cc++) x0[rr * 4 + cc] = guess ? (double)guess[cc * 4 + rr] : (rr == cc ? 1.0 : 0.0);
  r->lm_lambda = -1.0;
  r->converged = false;
  r->spec_ready = false;
  r->n_linearize = r->n_compute_error = 0;
  int iterations = 0;
  double last_y0 = 0.0;
  CK(c, cudaEventRecord(c->ev[6], c->stream));
  for (int i = 0; i < r->prm.max_iterations && !r->converged; i++) {
    iterations = i;
    double delta[16];
    int rc = (r->prm.optimizer == RGC_OPT_GAUSS_NEWTON) ? step_gn(r, x0, delta) : step_lm(r, x0, delta, &last_y0, i + 1 < r->prm.max_iterations);
    if (rc < 0) return rc;
    if (rc == 0) {
      std::fprintf(stderr, "lm not converged!!\n");  // lsq_registration_impl.hpp:69-72
      break;
    }
    r->converged = lm::is_converged(delta, r->prm.rotation_epsilon, r->prm.transformation_epsilon);
  }
  for (int i = 0; i < 16; i++) r->final_T[i] = (float)x0[i];  // final_transformation_ = x0.cast<float>()
  if (out_points) {
    RtF Tf;
    for (int i = 0; i < 12; i++) Tf.m[i] = r->final_T[i];
    Scratch tmp(c);
    float4* d_out = (float4*)tmp.get(sizeof(float4) * (size_t)r->src.n);
    if (!d_out) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (output cloud)");
    k_transform_out<<<div_up(r->src.n, 256), 256, 0, c->stream>>>(r->src.sorted, r->src.n, Tf, d_out);
    CKL(c);
    CK(c, cudaMemcpyAsync(out_points, d_out, sizeof(float4) * (size_t)r->src.n, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(c, cudaEventRecord(c->ev[7], c->stream));
  c->mark("align: end");
  CK(c, cudaEventSynchronize(c->ev[7]));
  c->print_marks();
  float total_ms = 0.f;
  CK(c, cudaEventElapsedTime(&total_ms, c->ev[5], c->ev[7]));
  CK(c, cudaEventElapsedTime(&r->lm_ms, c->ev[6], c->ev[7]));
  if (final_T16)
    for (int rr = 0; rr < 4; rr++)
      for (int cc = 0; cc < 4; cc++) final_T16[cc * 4 + rr] = r->final_T[rr * 4 + cc];
  if (result) {
    result->converged = r->converged ? 1 : 0;
    result->iterations = iterations;
    result->n_linearize = r->n_linearize;
    result->n_compute_error = r->n_compute_error;
    result->n_inliers = r->last_inliers;
    result->final_error = last_y0;
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) result->final_hessian[j * 6 + i] = r->final_hessian[i * 6 + j];
    result->device_ms = total_ms;
  }
  return RGC_OK;
}

int rgc_reg_linearize(rgc_reg* r, const double* T16, double* err, double* H36, double* b6) {
  if (!r || !T16 || !err) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_ready(r));
  double T[16], H[36];
  colmajor_to_row(T16, T);
  TRY(reg_linearize(r, T, err, (H36 && b6) ? H : nullptr, b6));
  if (H36 && b6)
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H36[j * 6 + i] = H[i * 6 + j];
  return RGC_OK;
}

int rgc_reg_compute_error(rgc_reg* r, const double* T16, double* err) {
  if (!r || !T16 || !err) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  TRY(join_side(c));
  double T[16];
  colmajor_to_row(T16, T);
  return reg_compute_error(r, T, err);
}

int rgc_reg_get_correspondences(rgc_reg* r, int32_t* corr, float* sq_dist) {
  if (!r || !corr) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  TRY(join_side(c));
  if (r->vgicp) FAIL(c, RGC_ERR_UNSUPPORTED, "point correspondences do not exist in voxelised mode");
  if (!r->have_corr) FAIL(c, RGC_ERR_STATE, "no correspondences yet (call linearize or align first)");
  const size_t n = (size_t)r->src.n;
  Scratch tmp(c);
  int* d_c = (int*)tmp.get(4 * n);
  float* d_s = (float*)tmp.get(4 * n);
  if (!d_c || !d_s) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (correspondence export)");
  k_corr_to_orig<<<div_up(r->src.n, 256), 256, 0, c->stream>>>(r->src.sorted, r->tgt.sorted, r->corr, r->sqd, r->src.n, d_c, d_s);
  CKL(c);
  CK(c, cudaMemcpyAsync(corr, d_c, 4 * n, cudaMemcpyDeviceToHost, c->stream));
  if (sq_dist) CK(c, cudaMemcpyAsync(sq_dist, d_s, 4 * n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

int rgc_reg_fitness(rgc_reg* r, double max_range, double* score) {
  if (!r || !score) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  TRY(join_side(c));
  if (!r->src.valid || !r->tgt.valid) FAIL(c, RGC_ERR_STATE, "source and target clouds must both be set");
  RtF Tf;
  for (int i = 0; i < 12; i++) Tf.m[i] = r->final_T[i];
  const int spread = query_spread(r->src.n);
  const int blocks = div_up(r->src.n * spread, kThreads);
  TRY(reg_ensure_fitness(r, blocks));
  const int* hint = (!r->vgicp && r->have_corr && r->corr) ? r->corr : nullptr;  // correspondences of the last linearize (same clouds)
  k_fitness<<<blocks, kThreads, 0, c->stream>>>(r->tgt.view, r->src.sorted, r->src.n, spread, Tf, max_range, r->slab, hint, r->fit_partials, c->d_ticket,
                                                reg_result_ptr(r), reg_next_done(r));
  CKL(c);
  TRY(reg_finish_reduce(r, 2));
  const double sum = c->h_result[0], nr = c->h_result[1];
  *score = nr > 0 ? sum / nr : DBL_MAX;
  return RGC_OK;
}

int rgc_reg_get_final_transformation(const rgc_reg* r, float* T16) {
  if (!r || !T16) return RGC_ERR_INVALID;
  for (int rr = 0; rr < 4; rr++)
    for (int cc = 0; cc < 4; cc++) T16[cc * 4 + rr] = r->final_T[rr * 4 + cc];
  return RGC_OK;
}

int rgc_knn(rgc_ctx* c, const void* points, size_t n, size_t stride, const void* queries, size_t m, size_t qstride, int k, int32_t* idx, float* d2,
            float grid_cell) {
  if (!c || !points || !queries || !idx || k < 1) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  if (k > kMaxK) FAIL(c, RGC_ERR_UNSUPPORTED, "k > 128 is not supported");
  if (qstride < 12 || qstride % 4) FAIL(c, RGC_ERR_INVALID, "query stride must be a multiple of 4 and >= 12 bytes");
  TmpCloud tc(c);
  Cloud& cl = tc.cl;
  TRY(cloud_build(c, cl, points, n, stride, false, 0, grid_cell));
  Scratch tmp(c);
  void* qraw = tmp.get(m * qstride);
  float4* q4 = (float4*)tmp.get(sizeof(float4) * m);
  float* bb = (float*)tmp.get(sizeof(float) * 6 * kBboxBlocks);
  int* d_idx = (int*)tmp.get(4 * m * (size_t)k);
  float* d_d2 = (float*)tmp.get(4 * m * (size_t)k);
  if (!qraw || !q4 || !bb || !d_idx || !d_d2) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (knn)");
  CK(c, cudaMemcpyAsync(qraw, queries, m * qstride, cudaMemcpyHostToDevice, c->stream));
  k_ingest<<<kBboxBlocks, 256, 0, c->stream>>>((const unsigned char*)qraw, qstride, (int)m, q4, bb);
  CKL(c);
  TRY(launch_knn<false>(c, cl.view, q4, (int)m, k, d_idx, d_d2));
  CK(c, cudaMemcpyAsync(idx, d_idx, 4 * m * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  if (d2) CK(c, cudaMemcpyAsync(d2, d_d2, 4 * m * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

// kNN of every point of a cloud within the cloud itself, through the SAME kernel calculate_covariances
// uses (test hook for fast_gicp_impl.hpp:254); idx: n x k original indices, ascending by (d2, index)
int rgc_knn_self(rgc_ctx* c, const void* points, size_t n, size_t stride, int k, int32_t* idx, float grid_cell) {
  if (!c || !points || !idx || k < 1) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  TmpCloud tc(c);
  Cloud& cl = tc.cl;
  TRY(cloud_build(c, cl, points, n, stride, false, 0, grid_cell));
  Scratch tmp(c);
  int* nbr = (int*)tmp.get(sizeof(int) * (size_t)k * n);
  int* d_idx = (int*)tmp.get(sizeof(int) * (size_t)k * n);
  if (!nbr || !d_idx) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (knn_self)");
  TRY(launch_knn_self(c, cl.view, cl.n, k, nbr));
  k_nbr_to_orig<<<div_up(cl.n, 256), 256, 0, c->stream>>>(cl.sorted, nbr, cl.n, k, d_idx);
  CKL(c);
  CK(c, cudaMemcpyAsync(idx, d_idx, sizeof(int) * (size_t)k * n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

// debug aid: candidate count above which the tile kNN hands a tile to the warp-per-query kernel
// (<= 0: never); the default comes from RGC_KNN_DEFER or 600
// test hook: how many cloud builds ran with a speculative grid geometry, and how many of those had to be redone
int rgc_debug_build_stats(rgc_ctx* c, unsigned long long* spec_builds, unsigned long long* spec_misses) {
  if (!c) return RGC_ERR_INVALID;
  if (spec_builds) *spec_builds = c->spec_builds;
  if (spec_misses) *spec_misses = c->spec_misses;
  return RGC_OK;
}
// builds whose level tables were filled with the previous cloud's sizes / how many of those had to be redone
int rgc_debug_table_stats(rgc_ctx* c, unsigned long long* spec_tables, unsigned long long* spec_table_misses) {
  if (!c) return RGC_ERR_INVALID;
  if (spec_tables) *spec_tables = c->spec_table_builds;
  if (spec_table_misses) *spec_table_misses = c->spec_table_misses;
  return RGC_OK;
}
int rgc_debug_set_knn_defer(rgc_ctx* c, int cands) {
  if (!c) return RGC_ERR_INVALID;
  c->knn_defer = cands;
  return RGC_OK;
}

// debug aid (not part of the documented ABI): per-warp statistics of the tile kNN kernel on a cloud;
// stats: ceil(n/32) x 4 int64 {cycles, nodes, candidates, fold steps}
int rgc_debug_tile_stats(rgc_ctx* c, const void* points, size_t n, size_t stride, int k, long long* stats, float grid_cell) {
  if (!c || !points || !stats) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  TmpCloud tc(c);
  Cloud& cl = tc.cl;
  TRY(cloud_build(c, cl, points, n, stride, false, 0, grid_cell));
  const size_t nw = (n + 31) / 32;
  Scratch tmp(c);
  int* nbr = (int*)tmp.get(sizeof(int) * (size_t)k * n);
  long long* d_stats = (long long*)tmp.get(sizeof(long long) * 4 * nw);
  if (!nbr || !d_stats) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (debug)");
  CK(c, cudaMemsetAsync(d_stats, 0, sizeof(long long) * 4 * nw, c->stream));
  CK(c, cudaMemcpyToSymbolAsync(g_tile_dbg, &d_stats, sizeof(d_stats), 0, cudaMemcpyHostToDevice, c->stream));
  TRY(launch_knn_self(c, cl.view, cl.n, k, nbr));
  long long* null_ptr = nullptr;
  CK(c, cudaMemcpyToSymbolAsync(g_tile_dbg, &null_ptr, sizeof(null_ptr), 0, cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaMemcpyAsync(stats, d_stats, sizeof(long long) * 4 * nw, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

// debug aid: per-query statistics of one correspondence search at pose T16 (column-major);
// stats: n_source x 4 int64 {cycles, nodes, table probes, candidates}, source in SORTED order
int rgc_debug_correspond_stats(rgc_reg* r, const double* T16, long long* stats) {
  if (!r || !T16 || !stats) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_ready(r));
  const size_t n = (size_t)r->src.n;
  Scratch tmp(c);
  long long* d_stats = (long long*)tmp.get(sizeof(long long) * 4 * n);
  if (!d_stats) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (debug)");
  CK(c, cudaMemsetAsync(d_stats, 0, sizeof(long long) * 4 * n, c->stream));
  CK(c, cudaMemcpyToSymbolAsync(g_tile_dbg, &d_stats, sizeof(d_stats), 0, cudaMemcpyHostToDevice, c->stream));
  double T[16], e;
  colmajor_to_row(T16, T);
  int rc = reg_linearize(r, T, &e, nullptr, nullptr);
  long long* null_ptr = nullptr;
  CK(c, cudaMemcpyToSymbolAsync(g_tile_dbg, &null_ptr, sizeof(null_ptr), 0, cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaMemcpyAsync(stats, d_stats, sizeof(long long) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return rc;
}

int rgc_reg_set_target_covariance_mode(rgc_reg* r, int on_demand) {
  if (!r) return RGC_ERR_INVALID;
  r->lazy_target = on_demand != 0;
  return RGC_OK;
}
int rgc_ctx_last_ondemand_ms(const rgc_ctx* c, float* ms) {
  if (!c || !ms) return RGC_ERR_INVALID;
  *ms = c->last_ondemand_ms;
  return RGC_OK;
}

// test hook (not part of the documented ABI): the target covariances AS THEY ARE on the device, without
// triggering any computation, in the caller's order; state[i] = 1 where covariance i has been computed
int rgc_debug_get_target_cov_state(rgc_reg* r, double* m4x4, int32_t* state) {
  if (!r || !m4x4 || !state) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  TRY(join_side(c));
  Cloud& t = r->tgt;
  if (!t.valid) FAIL(c, RGC_ERR_STATE, "no target cloud");
  const size_t n = (size_t)t.n;
  Scratch tmp(c);
  double* stage = (double*)tmp.get(128 * n);
  int* d_state = (int*)tmp.get(4 * n);
  if (!stage || !d_state) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (debug)");
  CK(c, cudaMemsetAsync(stage, 0, 128 * n, c->stream));
  if (t.cov) {
    k_cov_export<<<div_up(t.n, 256), 256, 0, c->stream>>>(t.sorted, t.n, t.cov, stage);
    CKL(c);
  }
  k_flags_to_orig<<<div_up(t.n, 256), 256, 0, c->stream>>>(t.sorted, t.n, (t.lazy_cov && !t.has_cov) ? t.cov_state : nullptr, t.has_cov ? 1 : 0, d_state);
  CKL(c);
  CK(c, cudaMemcpyAsync(m4x4, stage, 128 * n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(state, d_state, 4 * n, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

int rgc_reg_last_inliers(const rgc_reg* r, int* n) {
  if (!r || !n) return RGC_ERR_INVALID;
  *n = r->last_inliers;
  return RGC_OK;
}

int rgc_reg_set_vgicp(rgc_reg* r, int enabled, double resolution, int neighbor_search, int accumulation_mode) {
  if (!r) return RGC_ERR_INVALID;
  if (enabled && (!(resolution > 0.0) || neighbor_search < 0 || neighbor_search > 2 || accumulation_mode < 0 || accumulation_mode > 2))
    FAIL(r->ctx, RGC_ERR_INVALID, "bad voxel parameters");
  r->vgicp = enabled != 0;
  r->vox_res = resolution;
  r->vox_search = neighbor_search;
  r->vox_mode = accumulation_mode;
  r->vox_valid = false;
  r->have_corr = false;
  return RGC_OK;
}

int rgc_reg_get_voxels(rgc_reg* r, int32_t* coords3, int32_t* num_points, double* mean3, double* cov6, size_t cap, size_t* n_voxels) {
  if (!r || !n_voxels) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  TRY(reg_drain(r));
  TRY(join_side(c));
  if (!r->vgicp) FAIL(c, RGC_ERR_STATE, "voxelised mode is off (rgc_reg_set_vgicp)");
  if (!r->tgt.valid) FAIL(c, RGC_ERR_STATE, "no target cloud");
  TRY(cloud_covariances(c, r->tgt, r->prm.k_correspondences, r->prm.regularization));
  TRY(vgicp_build(r));
  *n_voxels = (size_t)r->vox_count;
  if (!coords3 || cap == 0) return RGC_OK;
  Scratch tmp(c);
  int *d_c = (int*)tmp.get(12 * cap), *d_n = (int*)tmp.get(4 * cap);
  double *d_m = (double*)tmp.get(24 * cap), *d_v = (double*)tmp.get(48 * cap);
  unsigned int* d_cnt = (unsigned int*)tmp.get(4);
  if (!d_c || !d_n || !d_m || !d_v || !d_cnt) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (voxel export)");
  CK(c, cudaMemsetAsync(d_cnt, 0, 4, c->stream));
  const uint32_t nslots = r->vox.mask + 1;
  k_vox_dump<<<div_up((int)nslots, 256), 256, 0, c->stream>>>(r->vox_slots, nslots, r->vox.geom, d_cnt, d_c, d_n, d_m, d_v, (unsigned)cap);
  CKL(c);
  const size_t m = std::min(cap, (size_t)r->vox_count);
  CK(c, cudaMemcpyAsync(coords3, d_c, 12 * m, cudaMemcpyDeviceToHost, c->stream));
  if (num_points) CK(c, cudaMemcpyAsync(num_points, d_n, 4 * m, cudaMemcpyDeviceToHost, c->stream));
  if (mean3) CK(c, cudaMemcpyAsync(mean3, d_m, 24 * m, cudaMemcpyDeviceToHost, c->stream));
  if (cov6) CK(c, cudaMemcpyAsync(cov6, d_v, 48 * m, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}

int rgc_reg_set_owner_slab(rgc_reg* r, int axis, float lo, float hi) {
  if (!r || axis > 2) return RGC_ERR_INVALID;
  r->slab = Slab{axis, lo, hi};
  r->have_corr = false;
  return RGC_OK;
}
// config C5: this rank's share of a large target, selected ON THE DEVICE from the full cloud: the points whose
// `axis` coordinate lies in [lo - halo, hi + halo) are kept (input order), become the target, and [lo, hi) becomes
// the ownership slab of rgc_reg_set_owner_slab
int rgc_reg_set_target_slab(rgc_reg* r, const void* pts, size_t n_sz, size_t stride, int axis, float lo, float hi, float halo, uint64_t key, size_t* n_local,
                            int32_t* local_index) {
  if (!r || !pts || axis < 0 || axis > 2 || !(halo >= 0.f)) return RGC_ERR_INVALID;
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  if (n_sz == 0 || n_sz > 0x7fffffffu) FAIL(c, RGC_ERR_INVALID, "target size out of range");
  if (stride < 12 || stride % 4) FAIL(c, RGC_ERR_INVALID, "point stride must be a multiple of 4 and >= 12 bytes");
  TRY(reg_drain(r));
  r->have_corr = false;
  r->vox_valid = false;
  r->slab = Slab{axis, lo, hi};
  const int n = (int)n_sz;
  cudaStream_t st = c->stream;
  Scratch tmp(c);
  const int nblk = div_up(n, 256 * kSlabItems);
  unsigned char* raw = (unsigned char*)tmp.get(n_sz * stride);
  unsigned int* blk = (unsigned int*)tmp.get(sizeof(unsigned int) * (size_t)(nblk + 1));
  if (!raw || !blk) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (slab staging)");
  CK(c, cudaMemcpyAsync(raw, pts, n_sz * stride, cudaMemcpyHostToDevice, st));
  const float klo = lo - halo, khi = hi + halo;
  k_slab_count<<<nblk, 256, 0, st>>>(raw, stride, n, axis, klo, khi, blk);
  CKL(c);
  k_vg_scan_blocks<<<1, 1024, 0, st>>>(blk, nblk);
  CKL(c);
  CK(c, cudaMemcpyAsync(c->h_counts, blk + nblk, 4, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  const size_t m = (size_t)c->h_counts[0];
  if (n_local) *n_local = m;
  if (m == 0) FAIL(c, RGC_ERR_INVALID, "no target point falls into this rank's slab + halo");
  float4* local = (float4*)tmp.get(sizeof(float4) * m);
  int* d_index = local_index ? (int*)tmp.get(sizeof(int) * m) : nullptr;
  if (!local || (local_index && !d_index)) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (slab)");
  k_slab_scatter<<<nblk, 256, 0, st>>>(raw, stride, n, axis, klo, khi, blk, local, d_index);
  CKL(c);
  if (local_index) CK(c, cudaMemcpyAsync(local_index, d_index, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
  TRY(cloud_build(c, r->tgt, local, m, sizeof(float4), true, key, r->prm.grid_cell));
  if (!target_lazy(r)) TRY(cloud_covariances(c, r->tgt, r->prm.k_correspondences, r->prm.regularization, true));
  if (local_index) CK(c, cudaStreamSynchronize(st));
  return RGC_OK;
}

int rgc_reg_set_comm(rgc_reg* r, rgc_comm* comm) {
  if (!r || (comm && comm->ctx != r->ctx)) return RGC_ERR_INVALID;
  r->comm = comm;
  return RGC_OK;
}
int rgc_reg_set_allreduce(rgc_reg* r, rgc_reduce_fn fn, void* user, void* d_buf) {
  if (!r || (fn && !d_buf)) return RGC_ERR_INVALID;
  r->reduce_fn = fn;
  r->reduce_user = user;
  r->reduce_buf = (double*)d_buf;
  return RGC_OK;
}

// ---- pre-step (include/rgc_preprocess.h) ----
static int pre_to_host(rgc_ctx* c, const void* pts, size_t n, size_t stride, size_t inten_off, float leaf, const double* q, const double* t, float period,
                       float* out_xyzi, size_t cap, size_t* n_out, int* passthrough) {
  if (!c || !n_out) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  float4* d = nullptr;
  int m = 0;
  TRY(pre_filter(c, pts, n, stride, inten_off, leaf, q, t, period, &d, &m, passthrough));
  Scratch tmp(c);
  tmp.blocks.push_back(d);  // handed over by pre_filter
  *n_out = (size_t)m;
  if (out_xyzi && cap > 0) CK(c, cudaMemcpyAsync(out_xyzi, d, sizeof(float4) * std::min((size_t)m, cap), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return RGC_OK;
}
int rgc_voxel_grid(rgc_ctx* c, const void* pts, size_t n, size_t stride, size_t inten_off, float leaf, float* out_xyzi, size_t cap, size_t* n_out, int* passthrough) {
  if (c && !(leaf > 0.f)) FAIL(c, RGC_ERR_INVALID, "leaf size must be positive");
  return pre_to_host(c, pts, n, stride, inten_off, leaf, nullptr, nullptr, 0.f, out_xyzi, cap, n_out, passthrough);
}
int rgc_deskew(rgc_ctx* c, const void* pts, size_t n, size_t stride, size_t inten_off, const double* q_wxyz, const double* t3, float scan_period, float* out_xyzi) {
  if (!q_wxyz || !t3 || !out_xyzi) return RGC_ERR_INVALID;
  size_t m = 0;
  return pre_to_host(c, pts, n, stride, inten_off, 0.f, q_wxyz, t3, scan_period, out_xyzi, n, &m, nullptr);
}
static int set_cloud_filtered(rgc_reg* r, Cloud& cl, const void* pts, size_t n, size_t stride, size_t inten_off, float leaf, const double* q, const double* t,
                              float period, uint64_t key, size_t* n_out) {
  rgc_ctx* c = r->ctx;
  CK(c, cudaSetDevice(c->device));
  if (key != 0 && cl.valid && cl.key == key) {
    if (n_out) *n_out = (size_t)cl.n;
    return RGC_OK;
  }
  r->have_corr = false;
  if (&cl == &r->tgt) r->vox_valid = false;
  SideLane side(c, &cl == &r->src);
  float4* d = nullptr;
  int m = 0;
  TRY(pre_filter(c, pts, n, stride, inten_off, leaf, q, t, period, &d, &m, nullptr));
  int rc = cloud_build(c, cl, d, (size_t)m, sizeof(float4), true, key, r->prm.grid_cell);
  c->put(d);  // same lane, stream-ordered reuse
  if (rc != RGC_OK) return rc;
  if (n_out) *n_out = (size_t)m;
  if (&cl == &r->tgt && target_lazy(r)) return RGC_OK;
  return cloud_covariances(c, cl, r->prm.k_correspondences, r->prm.regularization, true);
}
int rgc_reg_set_source_filtered(rgc_reg* r, const void* pts, size_t n, size_t stride, size_t inten_off, float leaf, const double* q_wxyz, const double* t3,
                                float scan_period, uint64_t key, size_t* n_out) {
  return r ? set_cloud_filtered(r, r->src, pts, n, stride, inten_off, leaf, q_wxyz, t3, scan_period, key, n_out) : RGC_ERR_INVALID;
}
int rgc_reg_set_target_filtered(rgc_reg* r, const void* pts, size_t n, size_t stride, size_t inten_off, float leaf, const double* q_wxyz, const double* t3,
                                float scan_period, uint64_t key, size_t* n_out) {
  return r ? set_cloud_filtered(r, r->tgt, pts, n, stride, inten_off, leaf, q_wxyz, t3, scan_period, key, n_out) : RGC_ERR_INVALID;
}

// ---- mapping-node association (include/rgc_mapping.h) ----
struct rgc_map {
  rgc_ctx* ctx = nullptr;
  Cloud cl;
};
int rgc_map_create(rgc_ctx* c, const void* pts, size_t n, size_t stride, rgc_map** out) {
  if (!c || !out) return RGC_ERR_INVALID;
  *out = nullptr;
  CK(c, cudaSetDevice(c->device));
  rgc_map* m = new rgc_map();
  m->ctx = c;
  int rc = cloud_build(c, m->cl, pts, n, stride, false, 0, 0.f);
  if (rc != RGC_OK) {
    cloud_release(c, m->cl);
    delete m;
    return rc;
  }
  *out = m;
  return RGC_OK;
}
int rgc_map_destroy(rgc_map* m) {
  if (!m) return RGC_OK;
  cudaSetDevice(m->ctx->device);
  cloud_release(m->ctx, m->cl);
  delete m;
  return RGC_OK;
}
}  // extern "C" (templates need C++ linkage)
template <bool PLANE>
static int map_associate(rgc_map* m, const void* feats, size_t n_sz, size_t stride, const double* q, const double* t, int32_t* valid, double* o1, double* o2,
                         size_t* n_valid) {
  if (!m || !feats || !q || !t || !valid || !o1 || !o2) return RGC_ERR_INVALID;
  rgc_ctx* c = m->ctx;
  CK(c, cudaSetDevice(c->device));
  if (n_sz == 0 || n_sz > 0x7fffffff / 32) FAIL(c, RGC_ERR_INVALID, "feature count out of range");
  if (stride < 12 || stride % 4) FAIL(c, RGC_ERR_INVALID, "point stride must be a multiple of 4 and >= 12 bytes");
  const int n = (int)n_sz;
  const size_t w2 = PLANE ? 1 : 3;
  Scratch tmp(c);
  void* staging = tmp.get(n_sz * stride);
  int* d_valid = (int*)tmp.get(4 * n_sz);
  double* d_o1 = (double*)tmp.get(24 * n_sz);
  double* d_o2 = (double*)tmp.get(8 * w2 * n_sz);
  if (!staging || !d_valid || !d_o1 || !d_o2) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (association)");
  cudaStream_t st = c->stream;
  CK(c, cudaMemcpyAsync(staging, feats, n_sz * stride, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemsetAsync(d_o1, 0, 24 * n_sz, st));
  CK(c, cudaMemsetAsync(d_o2, 0, 8 * w2 * n_sz, st));
  const PoseQ T{q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
  const int spread = query_spread(n);
  k_map_assoc<PLANE><<<div_up(n * spread, kThreads), kThreads, (size_t)5 * kThreads * 8, st>>>(m->cl.view, (const unsigned char*)staging, stride, n, spread, T, d_valid,
                                                                                          d_o1, d_o2);
  CKL(c);
  CK(c, cudaMemcpyAsync(valid, d_valid, 4 * n_sz, cudaMemcpyDeviceToHost, st));
  CK(c, cudaMemcpyAsync(o1, d_o1, 24 * n_sz, cudaMemcpyDeviceToHost, st));
  CK(c, cudaMemcpyAsync(o2, d_o2, 8 * w2 * n_sz, cudaMemcpyDeviceToHost, st));
  CK(c, cudaStreamSynchronize(st));
  if (n_valid) {
    size_t k = 0;
    for (size_t i = 0; i < n_sz; i++) k += valid[i] != 0;
    *n_valid = k;
  }
  return RGC_OK;
}
extern "C" {
int rgc_map_associate_edges(rgc_map* m, const void* feats, size_t n, size_t stride, const double* q, const double* t, int32_t* valid, double* pa, double* pb,
                            size_t* n_valid) {
  return map_associate<false>(m, feats, n, stride, q, t, valid, pa, pb, n_valid);
}
int rgc_map_associate_planes(rgc_map* m, const void* feats, size_t n, size_t stride, const double* q, const double* t, int32_t* valid, double* norm, double* dist,
                             size_t* n_valid) {
  return map_associate<true>(m, feats, n, stride, q, t, valid, norm, dist, n_valid);
}

int rgc_reg_stage_ms(const rgc_reg* r, float* ms7) {
  if (!r || !ms7) return RGC_ERR_INVALID;
  cloud_times(const_cast<Cloud&>(r->src));
  cloud_times(const_cast<Cloud&>(r->tgt));
  ms7[0] = r->src.build_ms;
  ms7[1] = r->src.knn_ms;
  ms7[2] = r->src.cov_ms;
  ms7[3] = r->tgt.build_ms;
  ms7[4] = r->tgt.knn_ms;
  ms7[5] = r->tgt.cov_ms;
  ms7[6] = r->lm_ms;
  return RGC_OK;
}

}  // extern "C"

#include "rgc_batch.inl"
