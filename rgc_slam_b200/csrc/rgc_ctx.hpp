// rgc_ctx.hpp — internal: the context object shared by the translation units of librgc_gicp.so
// (device, stream, pooled device memory, pinned result buffers) and the error-handling macros.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/rgc_gicp.h"

// ------------------------------------------------------------------------------------------------
struct rgc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  // pooled device memory: power-of-two size classes, never returned to the driver before destroy.
  // The context has two LANES (stream + block pool + pinned scratch): lane 0 is the main stream,
  // lane 1 carries the source-cloud preparation so that it overlaps the (much larger) target-cloud
  // preparation.  `stream`, `free_blocks`, `h_bbox`, `h_counts` always describe the CURRENT lane;
  // the other lane's copies are parked in `parked`.  A block always returns to the pool of the
  // lane that allocated it, so a pool's blocks are only ever reused in that lane's stream order.
  std::multimap<size_t, void*> free_blocks;
  struct BlockInfo {
    size_t size;
    int lane;
  };
  std::unordered_map<void*, BlockInfo> block_info;
  struct Parked {
    cudaStream_t stream = nullptr;
    std::multimap<size_t, void*> free_blocks;
    float* h_bbox = nullptr;
    uint32_t* h_counts = nullptr;
  } parked;
  int lane = 0;
  cudaEvent_t join_ev = nullptr;  // end of the last lane-1 work
  bool side_pending = false;      // lane-1 work not yet joined into the main stream
  bool overlap = std::getenv("RGC_NO_OVERLAP") == nullptr;
  bool look_ahead = std::getenv("RGC_NO_LOOKAHEAD") == nullptr;  // step_lm: linearize issued behind compute_error
  bool fuse_trial = std::getenv("RGC_NO_FUSE_TRIAL") == nullptr;  // compute_error inside the look-ahead correspondence launch (k_trial_step)
  bool late_join = std::getenv("RGC_NO_LATE_JOIN") == nullptr;    // main stream joins the source's build first, its covariances at the first k_linearize
  std::vector<cudaEvent_t> free_events;
  // small pinned slots (8 KB) for the host copies a cloud build waits on (bounding-box partials, per-level
  // cell counts): one per build in flight, so that deferred builds of different clouds never share one
  std::vector<float*> free_hslots;
  bool defer_builds = std::getenv("RGC_SYNC_BUILD") == nullptr;
  // grid geometry of the last cloud built on each lane: the next build on that lane generates its keys with it
  // before its own bounding box is known (rgc_gicp.cu: build_phase1 / build_phase3)
  struct GeomHint {
    bool valid = false;
    int nbits = 0;
    float s0 = 0.f, cell = 0.f;
    int cloud_bits = 0;
    // per-level table sizes of that cloud: the next build on the lane allocates, clears and FILLS its level
    // tables with them right behind the sort, while the per-level cell counts that used to size the tables
    // travel to the host beside it (build_phase3 checks them against these sizes)
    bool have_slots = false;
    uint32_t slots[20] = {};
    int slots_n = 0;  // points of the cloud those sizes were made for: a much larger cloud does not try them
  } geom_hint[2];
  bool spec_tables = std::getenv("RGC_NO_SPEC_TABLES") == nullptr;
  // programmatic dependent launch on the build / LM kernel chains (rgc_common.cuh: pdl_enter; rgc_gicp.cu: launch_pdl)
  bool pdl = std::getenv("RGC_NO_PDL") == nullptr;
  uint64_t spec_table_builds = 0, spec_table_misses = 0;
  cudaStream_t aux[2] = {nullptr, nullptr};  // per lane: the work a build forks off its lane (table clear, cell counts)
  bool spec_build = std::getenv("RGC_NO_SPEC_BUILD") == nullptr;
  uint64_t spec_builds = 0, spec_misses = 0;  // set_source / set_target return before the build's host waits
  // pinned, device-mapped result area the reduction kernels write straight into
  double* h_result = nullptr;
  double* d_result = nullptr;  // device alias of h_result
  float* h_bbox = nullptr;     // pinned: kBboxBlocks x 6
  uint32_t* h_counts = nullptr;  // pinned: kMaxLevels
  unsigned int* d_ticket = nullptr;
  // completion word of the reduction kernels (mapped pinned memory): the last block stores `seq` there
  // after its results are visible, and the host polls it instead of a cudaStreamSynchronize per LM step
  unsigned long long* h_seq = nullptr;
  unsigned long long* d_seq = nullptr;  // device alias of h_seq
  unsigned long long seq = 0;
  bool spin_wait = std::getenv("RGC_NO_SPIN") == nullptr;
  cudaEvent_t ev[8];
  cudaEvent_t evk[4];       // per-kernel timing of the last linearize / compute_error (profiling only)
  bool profile = false;     // rgc_ctx_set_profiling
  // candidate count at which a tile of the self-kNN is handed to the warp-per-query kernel; -1 = by cloud
  // size (rgc_debug_set_knn_defer / RGC_KNN_DEFER override)
  int knn_defer = std::getenv("RGC_KNN_DEFER") ? std::atoi(std::getenv("RGC_KNN_DEFER")) : -1;
  float last_kernel_ms[3] = {0, 0, 0};  // k_correspond, k_linearize, k_compute_error
  float last_ondemand_ms = 0;           // on-demand target kNN + covariances of the last linearize (profiling only)

  // RGC_TIMELINE=1: device timeline of one align (debug aid): events recorded behind the kernels of the build,
  // the covariance passes and the LM loop on whichever lane issues them, printed at the end of rgc_reg_align
  bool timeline = std::getenv("RGC_TIMELINE") != nullptr;
  struct Mark {
    const char* what;
    int lane;
    cudaEvent_t ev;
  };
  std::vector<Mark> marks;
  void mark(const char* what) {
    if (!timeline) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    marks.push_back(Mark{what, lane, e});
  }
  void print_marks() {
    if (!timeline || marks.empty()) return;
    cudaEventSynchronize(marks.back().ev);
    for (auto& m : marks) cudaEventSynchronize(m.ev);
    // sort by completion time relative to the first mark
    std::vector<std::pair<float, size_t>> order;
    for (size_t i = 0; i < marks.size(); i++) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[0].ev, marks[i].ev);
      order.push_back({ms, i});
    }
    std::sort(order.begin(), order.end());
    std::fprintf(stderr, "[rgc timeline] %zu marks (us since the first; lane 0 = main stream, 1 = source lane)\n", marks.size());
    float last[2] = {0.f, 0.f};
    for (auto& o : order) {
      const Mark& m = marks[o.second];
      std::fprintf(stderr, "[rgc timeline] %9.1f  (+%7.1f on lane %d)  %s\n", o.first * 1e3f, (o.first - last[m.lane]) * 1e3f, m.lane, m.what);
      last[m.lane] = o.first;
    }
    for (auto& m : marks) cudaEventDestroy(m.ev);
    marks.clear();
  }

  void switch_lane(int to) {
    if (to == lane) return;
    std::swap(stream, parked.stream);
    std::swap(free_blocks, parked.free_blocks);
    std::swap(h_bbox, parked.h_bbox);
    std::swap(h_counts, parked.h_counts);
    lane = to;
  }
  void* get(size_t bytes) {
    size_t cls = 4096;
    while (cls < bytes) cls <<= 1;
    auto it = free_blocks.find(cls);
    if (it != free_blocks.end()) {
      void* p = it->second;
      free_blocks.erase(it);
      return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, cls) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    block_info[p] = BlockInfo{cls, lane};
    return p;
  }
  void put(void* p) {
    if (!p) return;
    const BlockInfo bi = block_info[p];
    (bi.lane == lane ? free_blocks : parked.free_blocks).insert({bi.size, p});
  }
  cudaEvent_t get_event() {
    if (!free_events.empty()) {
      cudaEvent_t e = free_events.back();
      free_events.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return e;
  }
  void put_event(cudaEvent_t e) {
    if (e) free_events.push_back(e);
  }
  static constexpr size_t kHslotBytes = 8192;
  float* get_hslot() {
    if (!free_hslots.empty()) {
      float* p = free_hslots.back();
      free_hslots.pop_back();
      return p;
    }
    // mapped: the build kernels write the bounding-box partials and the level counts straight into it (a
    // cudaMemcpyAsync of a few bytes costs ~8 us of stream time)
    float* p = nullptr;
    if (cudaHostAlloc((void**)&p, kHslotBytes, cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return p;
  }
  static float* hslot_device(float* h) {
    float* d = nullptr;
    if (!h || cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return d;
  }
  void put_hslot(float* p) {
    if (p) free_hslots.push_back(p);
  }
};

// run the enclosed work on lane `to`, then return to the lane that was current
struct LaneScope {
  rgc_ctx* c;
  int prev;
  LaneScope(rgc_ctx* c_, int to) : c(c_), prev(c_->lane) { c->switch_lane(to); }
  ~LaneScope() { c->switch_lane(prev); }
};

// pooled scratch blocks of one call: every exit path (FAIL / CK / TRY included) gives them back
struct Scratch {
  rgc_ctx* c;
  std::vector<void*> blocks;
  explicit Scratch(rgc_ctx* c_) : c(c_) {}
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  void* get(size_t bytes) {
    void* p = c->get(bytes);
    if (p) blocks.push_back(p);
    return p;
  }
  // hand a block over to the caller (it is no longer returned to the pool by this object)
  void* release(void* p) {
    for (size_t i = 0; i < blocks.size(); i++)
      if (blocks[i] == p) {
        blocks.erase(blocks.begin() + (long)i);
        break;
      }
    return p;
  }
  ~Scratch() {
    for (void* p : blocks) c->put(p);
  }
};

#define CK(ctx, call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
      return RGC_ERR_CUDA;                                                                              \
    }                                                                                                   \
  } while (0)
#define CKL(ctx)                                                                                        \
  do {                                                                                                  \
    (ctx)->launches++;                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                                \
    if (e_ != cudaSuccess) {                                                                            \
      (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e_);                             \
      return RGC_ERR_CUDA;                                                                              \
    }                                                                                                   \
  } while (0)
#define FAIL(ctx, code, msg) \
  do {                       \
    (ctx)->err = (msg);      \
    return (code);           \
  } while (0)
#define TRY(expr)          \
  do {                     \
    int rc_ = (expr);      \
    if (rc_ != 0) return rc_; \
  } while (0)

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

