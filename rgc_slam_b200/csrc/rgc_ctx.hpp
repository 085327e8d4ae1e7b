// rgc_ctx.hpp — internal: the context object shared by the translation units of librgc_gicp.so
// (device, stream, pooled device memory, pinned result buffers) and the error-handling macros.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>

#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>

#include "../../include/rgc_gicp.h"

// ------------------------------------------------------------------------------------------------
struct rgc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  // pooled device memory: power-of-two size classes, never returned to the driver before destroy
  std::multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> block_size;
  // pinned, device-mapped result area the reduction kernels write straight into
  double* h_result = nullptr;
  double* d_result = nullptr;  // device alias of h_result
  float* h_bbox = nullptr;     // pinned: kBboxBlocks x 6
  uint32_t* h_counts = nullptr;  // pinned: kMaxLevels
  unsigned int* d_ticket = nullptr;
  cudaEvent_t ev[8];
  cudaEvent_t evk[4];       // per-kernel timing of the last linearize / compute_error (profiling only)
  bool profile = false;     // rgc_ctx_set_profiling
  int knn_defer = std::getenv("RGC_KNN_DEFER") ? std::atoi(std::getenv("RGC_KNN_DEFER")) : 600;  // rgc_debug_set_knn_defer
  float last_kernel_ms[3] = {0, 0, 0};  // k_correspond, k_linearize, k_compute_error

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
    block_size[p] = cls;
    return p;
  }
  void put(void* p) {
    if (!p) return;
    free_blocks.insert({block_size[p], p});
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

