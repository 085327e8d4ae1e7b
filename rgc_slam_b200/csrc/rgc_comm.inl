// rgc_comm.inl — NCCL communicator owned by the library (config C5: voxel/slab-sharded target); included by
// rgc_gicp.cu.  libnccl is opened at run time (dlopen), so the single-GPU library has no NCCL dependency
// and, inside a PyTorch process, the communicator uses the very libnccl.so.2 torch has already loaded.
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  const char* names[] = {std::getenv("RGC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.err = std::string("cannot open libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
    return &api;
  }
  auto sym = [&](const char* s) { return dlsym(api.handle, s); };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.err = "libnccl lacks the expected symbols";
  return &api;
}

}  // namespace

struct rgc_comm {
  rgc_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  double* d_buf = nullptr;  // 64 doubles: where the reduction kernels put a rank's partial sums
  uint64_t n_allreduce = 0;
  // all-reduce over peer memory (k_peer_allreduce): this rank's mailbox, every rank's mailbox as seen from here
  bool p2p = false;
  unsigned char* mailbox = nullptr;
  void* opened[kPeerMax] = {};
  PeerSet peers = {};
  unsigned long long p2p_seq = 0;
};
constexpr size_t kMailboxBytes = sizeof(unsigned long long) * 2 * kPeerMax * 2 * kPeerSlot;  // mail[2][kPeerMax][2 * kPeerSlot]

static void comm_p2p_release(rgc_comm* m) {
  for (int r = 0; r < kPeerMax; r++)
    if (m->opened[r]) {
      cudaIpcCloseMemHandle(m->opened[r]);
      m->opened[r] = nullptr;
    }
  if (m->mailbox) cudaFree(m->mailbox);
  m->mailbox = nullptr;
  m->p2p = false;
  cudaGetLastError();
}

// Map every rank's mailbox into this process (CUDA IPC; the handles travel through one ncclAllGather) — a collective,
// part of rgc_comm_create.  The ranks then agree (all-reduce of a success word) on whether ALL of them got there:
// the peer-memory all-reduce is used by everybody or by nobody (NCCL stays the transport otherwise, e.g. ranks on
// GPUs without peer access, or RGC_NO_P2P=1).
static int comm_p2p_setup(rgc_comm* m) {
  rgc_ctx* c = m->ctx;
  NcclApi* api = nccl_api();
  if (m->world < 2 || m->world > kPeerMax || !api->AllGather || std::getenv("RGC_NO_P2P")) return RGC_OK;
  bool ok = true;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  unsigned char* d_handles = nullptr;
  ok = ok && cudaMalloc((void**)&m->mailbox, kMailboxBytes) == cudaSuccess;
  ok = ok && cudaMemsetAsync(m->mailbox, 0, kMailboxBytes, c->stream) == cudaSuccess;  // ordered before the collectives below: every rank's
                                                                                        // mailbox is clear before any peer can write into it
  ok = ok && cudaIpcGetMemHandle(&mine, m->mailbox) == cudaSuccess;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  // the collectives below are entered by every rank whatever happened above
  if (cudaMalloc((void**)&d_handles, hb * (size_t)m->world) != cudaSuccess) {
    cudaGetLastError();
    FAIL(c, RGC_ERR_NOMEM, "device allocation failed (IPC handles)");
  }
  CK(c, cudaMemcpyAsync(d_handles + hb * (size_t)m->rank, &mine, hb, cudaMemcpyHostToDevice, c->stream));
  if (api->AllGather(d_handles + hb * (size_t)m->rank, d_handles, hb, ncclChar, m->comm, c->stream) != ncclSuccess) {
    cudaFree(d_handles);
    FAIL(c, RGC_ERR_CUDA, "ncclAllGather (IPC handles) failed");
  }
  std::vector<cudaIpcMemHandle_t> all((size_t)m->world);
  CK(c, cudaMemcpyAsync(all.data(), d_handles, hb * (size_t)m->world, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  cudaFree(d_handles);
  for (int r = 0; ok && r < m->world; r++) {
    void* p = m->mailbox;
    if (r != m->rank) {
      ok = cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      if (ok) m->opened[r] = p;
    }
    if (ok) m->peers.mail[r] = reinterpret_cast<unsigned long long*>(p);
  }
  cudaGetLastError();
  // everybody or nobody (this all-reduce is also the barrier behind every rank's cleared mailbox)
  const double mine_ok = ok ? 1.0 : 0.0;
  double all_ok = 0.0;
  CK(c, cudaMemcpyAsync(m->d_buf, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (api->AllReduce(m->d_buf, m->d_buf, 1, ncclDouble, ncclMin, m->comm, c->stream) != ncclSuccess) FAIL(c, RGC_ERR_CUDA, "ncclAllReduce (IPC agreement) failed");
  CK(c, cudaMemcpyAsync(&all_ok, m->d_buf, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  CK(c, cudaMemsetAsync(m->d_buf, 0, sizeof(double) * 64, c->stream));
  if (all_ok == 1.0)
    m->p2p = true;
  else
    comm_p2p_release(m);
  return RGC_OK;
}

// partial sums -> sum over ranks (in place, on the context's stream); with `publish`, the first n_a totals also go to
// dst_a, the next n_b to dst_b (the mapped host result area), followed by the completion word.  Over peer memory
// this is ONE launch; over NCCL the caller publishes with k_publish.
static int comm_allreduce(rgc_comm* m, int n_doubles, bool publish = false, int n_a = 0, double* dst_a = nullptr, int n_b = 0, double* dst_b = nullptr,
                          DoneFlag done = DoneFlag{nullptr, 0ull}) {
  rgc_ctx* c = m->ctx;
  if (m->p2p) {
    launch_pdl(c, k_peer_allreduce, dim3(1), dim3(64), 0, c->stream, m->d_buf, n_doubles, m->peers, m->rank, m->world, ++m->p2p_seq, publish ? n_a : 0,
               publish ? dst_a : nullptr, publish ? n_b : 0, publish ? dst_b : nullptr, done);
    CKL(c);
    m->n_allreduce++;
    return RGC_OK;
  }
  NcclApi* api = nccl_api();
  const ncclResult_t rc = api->AllReduce(m->d_buf, m->d_buf, (size_t)n_doubles, ncclDouble, ncclSum, m->comm, c->stream);
  if (rc != ncclSuccess) FAIL(c, RGC_ERR_CUDA, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error"));
  m->n_allreduce++;
  if (publish) {
    k_publish<<<1, 64, 0, c->stream>>>(m->d_buf, n_a, dst_a, n_b, dst_b, done);
    CKL(c);
  }
  return RGC_OK;
}

extern "C" {

int rgc_comm_unique_id(char* id128) {
  if (!id128) return RGC_ERR_INVALID;
  NcclApi* api = nccl_api();
  if (!api->err.empty()) {
    std::fprintf(stderr, "rgc_comm_unique_id: %s\n", api->err.c_str());
    return RGC_ERR_UNSUPPORTED;
  }
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return RGC_ERR_CUDA;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, 128);
  return RGC_OK;
}

int rgc_comm_create(rgc_ctx* c, const char* id128, int rank, int world, rgc_comm** out) {
  if (!c || !id128 || !out || world < 1 || rank < 0 || rank >= world) return RGC_ERR_INVALID;
  *out = nullptr;
  CK(c, cudaSetDevice(c->device));
  NcclApi* api = nccl_api();
  if (!api->err.empty()) FAIL(c, RGC_ERR_UNSUPPORTED, api->err);
  rgc_comm* m = new rgc_comm();
  m->ctx = c;
  m->rank = rank;
  m->world = world;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  const ncclResult_t rc = api->CommInitRank(&m->comm, world, id, rank);
  if (rc != ncclSuccess) {
    delete m;
    FAIL(c, RGC_ERR_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(rc) : "error"));
  }
  if (cudaMalloc((void**)&m->d_buf, sizeof(double) * 64) != cudaSuccess) {
    api->CommDestroy(m->comm);
    delete m;
    FAIL(c, RGC_ERR_NOMEM, "device allocation failed (communicator buffer)");
  }
  CK(c, cudaMemset(m->d_buf, 0, sizeof(double) * 64));
  // NCCL sets its channels up lazily at the first collective (~1 s): pay that here (rgc_comm_create is itself a
  // collective), not inside the first align
  for (int i = 0; i < 2; i++) {
    const ncclResult_t ar = api->AllReduce(m->d_buf, m->d_buf, 32, ncclDouble, ncclSum, m->comm, c->stream);
    if (ar != ncclSuccess) {
      api->CommDestroy(m->comm);
      cudaFree(m->d_buf);
      delete m;
      FAIL(c, RGC_ERR_CUDA, std::string("ncclAllReduce (warm-up): ") + (api->GetErrorString ? api->GetErrorString(ar) : "error"));
    }
  }
  CK(c, cudaStreamSynchronize(c->stream));
  {
    const int rc = comm_p2p_setup(m);
    if (rc != RGC_OK) {
      comm_p2p_release(m);
      api->CommDestroy(m->comm);
      cudaFree(m->d_buf);
      delete m;
      return rc;
    }
  }
  *out = m;
  return RGC_OK;
}

int rgc_comm_destroy(rgc_comm* m) {
  if (!m) return RGC_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  comm_p2p_release(m);
  if (m->comm) nccl_api()->CommDestroy(m->comm);
  cudaFree(m->d_buf);
  delete m;
  return RGC_OK;
}

// latency of the all-reduce the LM loop uses (n doubles, in place, on the context's stream): mean over `reps`
// back-to-back calls, CUDA events; a collective — every rank must call it
int rgc_comm_allreduce_us(rgc_comm* m, int n_doubles, int reps, float* us) {
  if (!m || !us || n_doubles < 1 || n_doubles > 64 || reps < 1) return RGC_ERR_INVALID;
  rgc_ctx* c = m->ctx;
  CK(c, cudaSetDevice(c->device));
  const uint64_t keep = m->n_allreduce;
  for (int i = 0; i < 3; i++) TRY(comm_allreduce(m, n_doubles));
  CK(c, cudaEventRecord(c->evk[0], c->stream));
  for (int i = 0; i < reps; i++) TRY(comm_allreduce(m, n_doubles));
  CK(c, cudaEventRecord(c->evk[1], c->stream));
  CK(c, cudaEventSynchronize(c->evk[1]));
  float ms = 0.f;
  CK(c, cudaEventElapsedTime(&ms, c->evk[0], c->evk[1]));
  *us = 1e3f * ms / (float)reps;
  m->n_allreduce = keep;
  CK(c, cudaMemsetAsync(m->d_buf, 0, sizeof(double) * 64, c->stream));
  return RGC_OK;
}

// Test hook: k_peer_allreduce with `world` ranks played by `world` streams of ONE device (every "rank" has its own mailbox
// and buffer in the same address space, so no IPC is needed): `reps` back-to-back all-reduces of n doubles per rank.
// in / out: [reps][world][n] host doubles; `published` (nullable): [reps][world][n], what each launch wrote through its
// dst_a (first n_a values) / dst_b (the rest) pointers.  The kernels of one repetition wait for each other, so all of a
// repetition's launches are issued before anything on the host blocks.
int rgc_debug_peer_allreduce_selftest(rgc_ctx* c, int world, int n, int n_a, int reps, const double* in, double* out, double* published) {
  if (!c || !in || !out || world < 1 || world > kPeerMax || n < 1 || n > kPeerSlot || n_a < 0 || n_a > n || reps < 1) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  const size_t per = (size_t)reps * world * kPeerSlot;
  unsigned char* boxes = nullptr;
  double *bufs = nullptr, *pub = nullptr;
  std::vector<cudaStream_t> st((size_t)world, nullptr);
  std::vector<double> h(per, 0.0);
  auto cleanup = [&]() {
    for (cudaStream_t s : st)
      if (s) cudaStreamDestroy(s);
    cudaFree(boxes);
    cudaFree(bufs);
    cudaFree(pub);
    cudaGetLastError();
  };
  bool ok = cudaMalloc((void**)&boxes, kMailboxBytes * (size_t)world) == cudaSuccess && cudaMalloc((void**)&bufs, sizeof(double) * per) == cudaSuccess &&
            cudaMalloc((void**)&pub, sizeof(double) * per) == cudaSuccess;
  for (int r = 0; ok && r < world; r++) ok = cudaStreamCreateWithFlags(&st[(size_t)r], cudaStreamNonBlocking) == cudaSuccess;
  if (!ok) {
    cleanup();
    FAIL(c, RGC_ERR_NOMEM, "allocation failed (peer all-reduce self-test)");
  }
  for (int rep = 0; rep < reps; rep++)
    for (int r = 0; r < world; r++)
      for (int t = 0; t < n; t++) h[((size_t)rep * world + r) * kPeerSlot + t] = in[((size_t)rep * world + r) * n + t];
  cudaMemset(boxes, 0, kMailboxBytes * (size_t)world);
  cudaMemset(pub, 0, sizeof(double) * per);
  cudaMemcpy(bufs, h.data(), sizeof(double) * per, cudaMemcpyHostToDevice);
  cudaDeviceSynchronize();
  PeerSet ps = {};
  for (int r = 0; r < world; r++) ps.mail[r] = reinterpret_cast<unsigned long long*>(boxes + kMailboxBytes * (size_t)r);
  for (int rep = 0; rep < reps; rep++)
    for (int r = 0; r < world; r++) {
      double* b = bufs + ((size_t)rep * world + r) * kPeerSlot;
      double* d = pub + ((size_t)rep * world + r) * kPeerSlot;
      k_peer_allreduce<<<1, 64, 0, st[(size_t)r]>>>(b, n, ps, r, world, (unsigned long long)(rep + 1), n_a, d, n - n_a, d + n_a, DoneFlag{nullptr, 0ull});
    }
  cudaError_t e = cudaGetLastError();
  for (int r = 0; r < world && e == cudaSuccess; r++) e = cudaStreamSynchronize(st[(size_t)r]);
  auto fetch = [&](const double* dev, double* host) {
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), dev, sizeof(double) * per, cudaMemcpyDeviceToHost);
    for (int rep = 0; rep < reps; rep++)
      for (int r = 0; r < world; r++)
        for (int t = 0; t < n; t++) host[((size_t)rep * world + r) * n + t] = h[((size_t)rep * world + r) * kPeerSlot + t];
  };
  fetch(bufs, out);
  if (published) fetch(pub, published);
  cleanup();
  CK(c, e);
  return RGC_OK;
}

int rgc_comm_transport(const rgc_comm* m) { return m ? (m->p2p ? 1 : 0) : RGC_ERR_INVALID; }

int rgc_comm_info(const rgc_comm* m, int* rank, int* world, uint64_t* n_allreduce, int* nccl_version) {
  if (!m) return RGC_ERR_INVALID;
  if (rank) *rank = m->rank;
  if (world) *world = m->world;
  if (n_allreduce) *n_allreduce = m->n_allreduce;
  if (nccl_version) {
    *nccl_version = 0;
    if (nccl_api()->GetVersion) nccl_api()->GetVersion(nccl_version);
  }
  return RGC_OK;
}

}  // extern "C"
