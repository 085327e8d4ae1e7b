// rgc_batch.inl — host side of rgc_batch_align (include/rgc_batch.h); included at the end of rgc_gicp.cu.
//
// One chunk = B pairs: two multi-cloud grids (all sources / all targets), the sources' k-NN + covariances in
// one launch, then LM rounds.  A round evaluates, for every pair still iterating, compute_error at the trial
// pose and the linearization at that same pose (the look-ahead of step_lm in rgc_gicp.cu: if the trial is
// accepted it IS the next iteration's linearization, if not it is discarded), so every round is the same
// five launches whatever state the pairs are in.  The per-pair step control below is step_lm / step_gn of
// rgc_gicp.cu (lsq_registration_impl.hpp:106-172) unrolled into a state machine.

namespace {

struct PairLM {
  double x0[16];
  double H[36], b[6], y0 = 0.0;
  double xi[16], delta[16], d[6];
  double lambda = -1.0, nu = 2.0;
  int inner = 0;  // trials made in the current outer iteration
  int outer = 0;  // index of the current outer iteration (nr_iterations_)
  int cur = 0;    // buffer set holding the correspondences / Mahalanobis matrices of the current linearization
  bool active = true, converged = false;
  int n_lin = 0, n_ce = 0, inliers = 0;
  double final_H[36];
  double last_y0 = 0.0;
};

struct BatchStats {
  float ms[6] = {0, 0, 0, 0, 0, 0};
  int rounds = 0;
};
static thread_local BatchStats g_batch_stats;

static void unpack_lin(const double* res, PairLM& p) {
  p.y0 = res[0];
  p.inliers = (int)res[kAccN];
  int o = 1;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) {
      p.H[i * 6 + j] = p.H[j * 6 + i] = res[o];
      o++;
    }
  for (int i = 0; i < 6; i++) p.b[i] = res[22 + i];
}

// pinned host scratch that lives as long as the call
struct Pinned {
  void* p = nullptr;
  ~Pinned() {
    if (p) cudaFreeHost(p);
  }
  bool alloc(size_t bytes) { return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess; }
};

// the inputs of one chunk on the device: raw records uploaded and converted on the side lane (so that the
// upload of chunk i + 1 overlaps the LM rounds of chunk i), then handed to the main stream through `ready`
struct ChunkIn {
  Scratch tmp;
  int B = 0, S = 0, Tn = 0;
  std::vector<int> soff, toff;
  float4 *pts_s = nullptr, *pts_t = nullptr;
  cudaEvent_t begin = nullptr, ready = nullptr;
  rgc_ctx* c;
  explicit ChunkIn(rgc_ctx* c_) : tmp(c_), c(c_) {}
  ~ChunkIn() {
    c->put_event(begin);
    c->put_event(ready);
  }
};

static int chunk_upload(rgc_ctx* c, const rgc_pair* pairs, int B, std::unique_ptr<ChunkIn>& out) {
  out.reset(new ChunkIn(c));
  ChunkIn& in = *out;
  LaneScope lane(c, c->overlap ? 1 : 0);
  cudaStream_t st = c->stream;
  in.B = B;
  in.soff.assign(B + 1, 0);
  in.toff.assign(B + 1, 0);
  std::vector<BCloudSrc> scs(B), tcs(B);
  size_t sbytes = 0, tbytes = 0;
  for (int p = 0; p < B; p++) {
    const rgc_pair& q = pairs[p];
    if (!q.source || !q.target || q.n_source == 0 || q.n_target == 0) FAIL(c, RGC_ERR_INVALID, "empty point cloud in the batch");
    if (q.source_stride < 12 || q.source_stride % 4 || q.target_stride < 12 || q.target_stride % 4)
      FAIL(c, RGC_ERR_INVALID, "point stride must be a multiple of 4 and >= 12 bytes");
    if ((size_t)in.soff[p] + q.n_source > 0x7fffffffu / 32 || (size_t)in.toff[p] + q.n_target > 0x7fffffffu / 32)
      FAIL(c, RGC_ERR_UNSUPPORTED, "chunk too large for 32-bit indexing");
    in.soff[p + 1] = in.soff[p] + (int)q.n_source;
    in.toff[p + 1] = in.toff[p] + (int)q.n_target;
    scs[p] = BCloudSrc{(unsigned long long)sbytes, (unsigned)q.source_stride, 0u};
    tcs[p] = BCloudSrc{(unsigned long long)tbytes, (unsigned)q.target_stride, 0u};
    sbytes += q.n_source * q.source_stride;
    tbytes += q.n_target * q.target_stride;
  }
  const int S = in.S = in.soff[B], Tn = in.Tn = in.toff[B];
  in.begin = c->get_event();
  in.ready = c->get_event();
  if (!in.begin || !in.ready) FAIL(c, RGC_ERR_CUDA, "cudaEventCreate failed");
  CK(c, cudaEventRecord(in.begin, st));
  unsigned char* raw_s = (unsigned char*)in.tmp.get(sbytes);
  unsigned char* raw_t = (unsigned char*)in.tmp.get(tbytes);
  in.pts_s = (float4*)in.tmp.get(sizeof(float4) * (size_t)S);
  in.pts_t = (float4*)in.tmp.get(sizeof(float4) * (size_t)Tn);
  BCloudSrc* d_scs = (BCloudSrc*)in.tmp.get(sizeof(BCloudSrc) * (size_t)B);
  BCloudSrc* d_tcs = (BCloudSrc*)in.tmp.get(sizeof(BCloudSrc) * (size_t)B);
  int* d_soff = (int*)in.tmp.get(sizeof(int) * (size_t)(B + 1));
  int* d_toff = (int*)in.tmp.get(sizeof(int) * (size_t)(B + 1));
  if (!raw_s || !raw_t || !in.pts_s || !in.pts_t || !d_scs || !d_tcs || !d_soff || !d_toff) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (batch staging)");
  for (int p = 0; p < B; p++) {
    CK(c, cudaMemcpyAsync(raw_s + scs[p].byte_off, pairs[p].source, pairs[p].n_source * pairs[p].source_stride, cudaMemcpyHostToDevice, st));
    CK(c, cudaMemcpyAsync(raw_t + tcs[p].byte_off, pairs[p].target, pairs[p].n_target * pairs[p].target_stride, cudaMemcpyHostToDevice, st));
  }
  CK(c, cudaMemcpyAsync(d_scs, scs.data(), sizeof(BCloudSrc) * (size_t)B, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(d_tcs, tcs.data(), sizeof(BCloudSrc) * (size_t)B, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(d_soff, in.soff.data(), sizeof(int) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(d_toff, in.toff.data(), sizeof(int) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
  k_bingest<<<div_up(S, 256), 256, 0, st>>>(raw_s, d_scs, d_soff, B, S, in.pts_s);
  CKL(c);
  k_bingest<<<div_up(Tn, 256), 256, 0, st>>>(raw_t, d_tcs, d_toff, B, Tn, in.pts_t);
  CKL(c);
  CK(c, cudaEventRecord(in.ready, st));
  return RGC_OK;
}

static int batch_chunk(rgc_ctx* c, const rgc_params& prm, const rgc_pair* pairs, ChunkIn& in, int want_fitness, double max_range, bool lazy, rgc_pair_result* out) {
  cudaStream_t st = c->stream;
  cudaEvent_t ev[8];
  for (int i = 0; i < 8; i++) ev[i] = c->ev[i];
  const int k = prm.k_correspondences;
  const int B = in.B, S = in.S, Tn = in.Tn;
  const std::vector<int>&soff = in.soff, &toff = in.toff;
  float4 *pts_s = in.pts_s, *pts_t = in.pts_t;
  Scratch tmp(c);
  CK(c, cudaStreamWaitEvent(st, in.ready, 0));
  CK(c, cudaEventRecord(ev[1], st));

  // ---- the two multi-cloud grids
  TmpCloud srcs_h(c), tgts_h(c);
  Cloud &srcs = srcs_h.cl, &tgts = tgts_h.cl;
  TRY(cloud_build(c, srcs, pts_s, (size_t)S, sizeof(float4), true, 0, prm.grid_cell, soff.data(), B));
  CK(c, cudaEventRecord(ev[2], st));
  TRY(cloud_build(c, tgts, pts_t, (size_t)Tn, sizeof(float4), true, 0, prm.grid_cell, toff.data(), B));
  CK(c, cudaEventRecord(ev[3], st));

  // ---- covariances: all sources now; targets on demand (or now, in the reference's schedule)
  TRY(cloud_tiles(c, srcs));
  TRY(cloud_covariances(c, srcs, k, prm.regularization));
  if (lazy) {
    tgts.cov = (double*)c->get(sizeof(double) * 6 * (size_t)Tn);
    tgts.cov_state = (int*)c->get(sizeof(int) * (size_t)Tn);
    if (!tgts.cov || !tgts.cov_state) FAIL(c, RGC_ERR_NOMEM, "device allocation failed (on-demand covariances)");
    CK(c, cudaMemsetAsync(tgts.cov_state, 0, sizeof(int) * (size_t)Tn, st));
    tgts.cov_k = k;
    tgts.cov_method = prm.regularization;
    tgts.lazy_cov = true;
  } else {
    TRY(cloud_tiles(c, tgts));
    TRY(cloud_covariances(c, tgts, k, prm.regularization));
  }
  CK(c, cudaEventRecord(ev[4], st));

  // ---- per-pair launch geometry
  std::vector<BPairInfo> info(B);
  std::vector<int> blk_pair, fblk_pair;
  int min_tgt = INT_MAX;
  for (int p = 0; p < B; p++) {
    BPairInfo& pi = info[p];
    const int ns = soff[p + 1] - soff[p];
    pi.src_lo = soff[p];
    pi.src_hi = soff[p + 1];
    pi.tgt_lo = toff[p];
    pi.tgt_hi = toff[p + 1];
    pi.tgt_prefix = (unsigned long long)p << (3 * tgts.view.nbits);
    pi.blk0 = (int)blk_pair.size();
    pi.nblk = reduce_grid(ns);
    pi.spread = query_spread(ns);
    pi.fblk0 = (int)fblk_pair.size();
    pi.fnblk = div_up(ns * pi.spread, kThreads);
    pi.pad = 0;
    blk_pair.insert(blk_pair.end(), (size_t)pi.nblk, p);
    if (want_fitness) fblk_pair.insert(fblk_pair.end(), (size_t)pi.fnblk, p);
    min_tgt = std::min(min_tgt, pi.tgt_hi - pi.tgt_lo);
  }
  const int nblk = (int)blk_pair.size(), nfblk = (int)fblk_pair.size();

  BPairInfo* d_info = (BPairInfo*)tmp.get(sizeof(BPairInfo) * (size_t)B);
  int* d_blk_pair = (int*)tmp.get(sizeof(int) * (size_t)nblk);
  int* d_fblk_pair = (int*)tmp.get(sizeof(int) * (size_t)std::max(nfblk, 1));
  BPairRound* d_rounds = (BPairRound*)tmp.get(sizeof(BPairRound) * (size_t)B);
  int* corr[2];
  float* sqd[2];
  double* maha[2];
  for (int s = 0; s < 2; s++) {
    corr[s] = (int*)tmp.get(4 * (size_t)S);
    sqd[s] = (float*)tmp.get(4 * (size_t)S);
    maha[s] = (double*)tmp.get(48 * (size_t)S);
  }
  int* need_list = (int*)tmp.get(4 * (size_t)S);
  int* need_nbr = (int*)tmp.get(4 * (size_t)S * (size_t)k);
  int* need_count = (int*)tmp.get(4);
  double* lin_partials = (double*)tmp.get(sizeof(double) * kLinN * (size_t)nblk);
  double* ce_partials = (double*)tmp.get(sizeof(double) * (size_t)nblk);
  double* fit_partials = (double*)tmp.get(sizeof(double) * 2 * (size_t)std::max(nfblk, 1));
  unsigned int* tickets = (unsigned int*)tmp.get(sizeof(unsigned int) * 3 * (size_t)B);  // linearize | compute_error | fitness
  double* d_res = (double*)tmp.get(sizeof(double) * 36 * (size_t)B);                     // per pair: 32 linearize | then B compute_error | 2B fitness
  float* d_final = (float*)tmp.get(sizeof(float) * 16 * (size_t)B);
  if (!d_info || !d_blk_pair || !d_fblk_pair || !d_rounds || !corr[1] || !sqd[1] || !maha[1] || !corr[0] || !sqd[0] || !maha[0] || !need_list || !need_nbr ||
      !need_count || !lin_partials || !ce_partials || !fit_partials || !tickets || !d_res || !d_final)
    FAIL(c, RGC_ERR_NOMEM, "device allocation failed (batch work buffers)");
  double* d_lin_res = d_res;                       // B x 32
  double* d_ce_res = d_res + 32 * (size_t)B;       // B
  double* d_fit_res = d_res + 33 * (size_t)B;      // B x 2
  Pinned pin;
  const size_t pin_bytes = sizeof(BPairRound) * (size_t)B + sizeof(double) * 36 * (size_t)B + sizeof(float) * 16 * (size_t)B;
  if (!pin.alloc(pin_bytes)) FAIL(c, RGC_ERR_NOMEM, "pinned allocation failed (batch)");
  BPairRound* h_rounds = (BPairRound*)pin.p;
  double* h_res = (double*)(h_rounds + B);
  float* h_final = (float*)(h_res + 36 * (size_t)B);
  CK(c, cudaMemcpyAsync(d_info, info.data(), sizeof(BPairInfo) * (size_t)B, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemcpyAsync(d_blk_pair, blk_pair.data(), sizeof(int) * (size_t)nblk, cudaMemcpyHostToDevice, st));
  if (nfblk) CK(c, cudaMemcpyAsync(d_fblk_pair, fblk_pair.data(), sizeof(int) * (size_t)nfblk, cudaMemcpyHostToDevice, st));
  CK(c, cudaMemsetAsync(tickets, 0, sizeof(unsigned int) * 3 * (size_t)B, st));
  CK(c, cudaMemsetAsync(need_count, 0, 4, st));

  // ---- LM state
  std::vector<PairLM> lm((size_t)B);
  for (int p = 0; p < B; p++) {
    PairLM& s = lm[p];
    for (int rr = 0; rr < 4; rr++)
      for (int cc = 0; cc < 4; cc++) s.x0[rr * 4 + cc] = (double)pairs[p].guess[cc * 4 + rr];  // lsq_registration_impl.hpp:54
    for (int i = 0; i < 36; i++) s.final_H[i] = (i % 7 == 0) ? 1.0 : 0.0;                      // final_hessian_.setIdentity()
    s.active = prm.max_iterations > 0;
  }
  const float thr = prm.max_correspondence_distance;
  const float thr2 = thr * thr;
  const bool gn = prm.optimizer == RGC_OPT_GAUSS_NEWTON;

  // one round: (compute_error at T) + correspondences + on-demand target covariances + linearize, all pairs
  auto launch_round = [&](bool any_ce) -> int {
    CK(c, cudaMemcpyAsync(d_rounds, h_rounds, sizeof(BPairRound) * (size_t)B, cudaMemcpyHostToDevice, st));
    if (any_ce) {
      k_bcompute_error<<<nblk, kThreads, 0, st>>>(tgts.sorted, srcs.sorted, d_info, d_blk_pair, d_rounds, corr[0], corr[1], maha[0], maha[1], ce_partials, tickets + B,
                                                  d_ce_res);
      CKL(c);
    }
    k_bcorrespond<<<nblk, kThreads, 0, st>>>(tgts.view, srcs.sorted, d_info, d_blk_pair, d_rounds, thr2, corr[0], corr[1], sqd[0], sqd[1],
                                             lazy ? tgts.cov_state : nullptr, need_list, need_count);
    CKL(c);
    if (lazy) {
      k_knn_warp<<<std::min(div_up(S, KW_WARPS), 148 * 8), KW_WARPS * 32, 0, st>>>(tgts.view, Tn, k, need_count, nullptr, need_list, S, need_nbr, nullptr, tgts.d_off, B);
      CKL(c);
      TRY(launch_covariance(c, st, tgts.sorted, need_nbr, S, S, k, min_tgt >= k, prm.regularization, tgts.cov, need_list, need_count));
      CK(c, cudaMemsetAsync(need_count, 0, 4, st));
    }
    k_blinearize<<<nblk, kThreads, 0, st>>>(tgts.sorted, srcs.sorted, srcs.cov, tgts.cov, d_info, d_blk_pair, d_rounds, corr[0], corr[1], maha[0], maha[1], lin_partials,
                                            tickets, d_lin_res);
    CKL(c);
    CK(c, cudaMemcpyAsync(h_res, d_res, sizeof(double) * 33 * (size_t)B, cudaMemcpyDeviceToHost, st));
    CK(c, cudaStreamSynchronize(st));
    g_batch_stats.rounds++;
    return RGC_OK;
  };
  auto set_round = [&](int p, const double* T, int do_ce, int rsel, int wsel, int hint_sel) {
    BPairRound& r = h_rounds[p];
    for (int i = 0; i < 12; i++) {
      r.T[i] = T[i];
      r.Tf[i] = (float)T[i];
    }
    r.active = 1;
    r.do_ce = do_ce;
    r.rsel = rsel;
    r.wsel = wsel;
    r.hint_sel = hint_sel;
    r.want_hb = 1;
  };
  auto finish = [&](PairLM& s) { s.active = false; };

  // round 0: linearize(x0) of every pair
  bool any = false;
  for (int p = 0; p < B; p++) {
    h_rounds[p].active = 0;
    if (!lm[p].active) continue;
    set_round(p, lm[p].x0, 0, 0, 0, -1);
    any = true;
  }
  if (any) {
    TRY(launch_round(false));
    for (int p = 0; p < B; p++) {
      if (!lm[p].active) continue;
      unpack_lin(h_res + 32 * (size_t)p, lm[p]);
      lm[p].n_lin = 1;
      lm[p].cur = 0;
      if (!gn) lm[p].last_y0 = lm[p].y0;
    }
  }
  while (any) {
    any = false;
    // ---- host: the next trial of every active pair
    for (int p = 0; p < B; p++) {
      PairLM& s = lm[p];
      h_rounds[p].active = 0;
      if (!s.active) continue;
      if (gn) {
        // step_gn (lsq_registration_impl.hpp:106-122): solve, move, then (next round) linearize at the new pose
        double nb[6];
        for (int i = 0; i < 6; i++) nb[i] = -s.b[i];
        lm::solve_ldlt6(s.H, nb, s.d);
        lm::se3_delta(s.d, s.delta);
        lm::mul4(s.delta, s.x0, s.x0);
        std::memcpy(s.final_H, s.H, sizeof(s.H));
        s.converged = lm::is_converged(s.delta, prm.rotation_epsilon, prm.transformation_epsilon);
        if (s.converged || s.outer + 1 >= prm.max_iterations) {
          finish(s);
          continue;
        }
        s.outer++;
        set_round(p, s.x0, 0, 0, 0, 0);
        any = true;
        continue;
      }
      if (s.lambda < 0.0) {  // lsq_registration_impl.hpp:130-132
        double mx = 0.0;
        for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(s.H[i * 7]));
        s.lambda = prm.lm_init_lambda_factor * mx;
      }
      double A[36], nb[6];
      for (int j = 0; j < 36; j++) A[j] = s.H[j];
      for (int j = 0; j < 6; j++) {
        A[j * 7] += s.lambda;
        nb[j] = -s.b[j];
      }
      lm::solve_ldlt6(A, nb, s.d);
      lm::se3_delta(s.d, s.delta);
      lm::mul4(s.delta, s.x0, s.xi);
      set_round(p, s.xi, 1, s.cur, 1 - s.cur, s.cur);
      any = true;
    }
    if (!any) break;
    TRY(launch_round(!gn));
    // ---- host: accept / reject (lsq_registration_impl.hpp:144-171)
    any = false;
    for (int p = 0; p < B; p++) {
      PairLM& s = lm[p];
      if (!s.active) continue;
      if (gn) {
        unpack_lin(h_res + 32 * (size_t)p, s);
        s.n_lin++;
        any = true;
        continue;
      }
      s.n_ce++;
      const double yi = h_res[32 * (size_t)B + p];
      double denom = 0.0;
      for (int j = 0; j < 6; j++) denom += s.d[j] * (s.lambda * s.d[j] - s.b[j]);
      const double rho = (s.y0 - yi) / denom;
      if (rho < 0) {
        if (lm::is_converged(s.delta, prm.rotation_epsilon, prm.transformation_epsilon)) {
          s.converged = true;  // step_lm returns true without moving; the outer loop then sees a converged delta
          finish(s);
          continue;
        }
        s.lambda = s.nu * s.lambda;
        s.nu = 2 * s.nu;
        if (++s.inner >= prm.lm_max_iterations) {
          std::fprintf(stderr, "lm not converged!!\n");  // lsq_registration_impl.hpp:69-72
          finish(s);
          continue;
        }
        any = true;
        continue;
      }
      std::memcpy(s.x0, s.xi, sizeof(s.xi));
      s.lambda = s.lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      std::memcpy(s.final_H, s.H, sizeof(s.H));
      s.converged = lm::is_converged(s.delta, prm.rotation_epsilon, prm.transformation_epsilon);
      if (s.converged || s.outer + 1 >= prm.max_iterations) {
        finish(s);
        continue;
      }
      // the linearization at xi computed in this round becomes the current one
      s.outer++;
      unpack_lin(h_res + 32 * (size_t)p, s);
      s.cur = 1 - s.cur;
      s.n_lin++;
      s.nu = 2.0;
      s.inner = 0;
      s.last_y0 = s.y0;
      any = true;
    }
  }
  CK(c, cudaEventRecord(ev[5], st));

  // ---- getFitnessScore at the final transformations
  for (int p = 0; p < B; p++) {
    for (int i = 0; i < 12; i++) h_final[p * 16 + i] = (float)lm[p].x0[i];  // final_transformation_ = x0.cast<float>()
    const int hsel = lm[p].n_lin > 0 ? lm[p].cur : -1;
    std::memcpy(&h_final[p * 16 + 12], &hsel, sizeof(int));
  }
  if (want_fitness) {
    CK(c, cudaMemcpyAsync(d_final, h_final, sizeof(float) * 16 * (size_t)B, cudaMemcpyHostToDevice, st));
    k_bfitness_search<<<nblk, kThreads, 0, st>>>(tgts.view, srcs.sorted, d_info, d_blk_pair, d_final, corr[0], corr[1], sqd[0], sqd[1]);
    CKL(c);
    k_bfitness<<<nfblk, kThreads, 0, st>>>(d_info, d_fblk_pair, d_final, max_range, sqd[0], sqd[1], fit_partials, tickets + 2 * B, d_fit_res);
    CKL(c);
    CK(c, cudaMemcpyAsync(h_res + 33 * (size_t)B, d_fit_res, sizeof(double) * 2 * (size_t)B, cudaMemcpyDeviceToHost, st));
  }
  CK(c, cudaEventRecord(ev[6], st));
  CK(c, cudaEventSynchronize(ev[6]));
  float ms[6];
  CK(c, cudaEventElapsedTime(&ms[0], in.begin, in.ready));  // on the side lane: overlaps the previous chunk's rounds
  for (int i = 1; i < 6; i++) CK(c, cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
  float total = 0.f;
  for (int i = 0; i < 6; i++) {
    g_batch_stats.ms[i] += ms[i];
    total += ms[i];
  }

  // ---- results
  for (int p = 0; p < B; p++) {
    const PairLM& s = lm[p];
    rgc_pair_result& o = out[p];
    for (int rr = 0; rr < 4; rr++)
      for (int cc = 0; cc < 4; cc++) o.final_T[cc * 4 + rr] = (float)s.x0[rr * 4 + cc];
    o.result.converged = s.converged ? 1 : 0;
    o.result.iterations = s.outer;
    o.result.n_linearize = s.n_lin;
    o.result.n_compute_error = s.n_ce;
    o.result.n_inliers = s.inliers;
    o.result.final_error = s.last_y0;
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) o.result.final_hessian[j * 6 + i] = s.final_H[i * 6 + j];
    o.result.device_ms = total / (float)B;
    o.fitness = 0.0;
    if (want_fitness) {
      const double sum = h_res[33 * (size_t)B + 2 * (size_t)p], nr = h_res[33 * (size_t)B + 2 * (size_t)p + 1];
      o.fitness = nr > 0 ? sum / nr : DBL_MAX;
    }
  }
  return RGC_OK;
}

}  // namespace

extern "C" {

int rgc_batch_align(rgc_ctx* c, const rgc_params* prm_in, const rgc_pair* pairs, size_t n_pairs, int want_fitness, double fitness_max_range, int max_chunk_pairs,
                    rgc_pair_result* out) {
  if (!c || (n_pairs && (!pairs || !out))) return RGC_ERR_INVALID;
  CK(c, cudaSetDevice(c->device));
  rgc_params prm;
  if (prm_in)
    prm = *prm_in;
  else
    rgc_params_default(&prm);
  if (prm.k_correspondences < 1 || prm.k_correspondences > 32) FAIL(c, RGC_ERR_UNSUPPORTED, "k_correspondences must be in [1, 32]");
  if (prm.regularization < 0 || prm.regularization > 4) FAIL(c, RGC_ERR_INVALID, "unknown regularization method");
  TRY(join_side(c));
  g_batch_stats = BatchStats();
  const bool lazy = std::getenv("RGC_EAGER_TARGET_COV") == nullptr;
  const size_t max_points = 24u << 20;
  // chunk boundaries first, then a two-stage pipeline: the upload + ingest of chunk i + 1 is queued on the side
  // lane before the (host-synchronous) LM rounds of chunk i start on the main stream
  // (Ramping the first chunks up from 1/8 of the capacity, so that less of the first upload goes unoverlapped, was
  // measured and is neutral: 22 402 vs 22 411 pairs/s on 8 GPUs, 6 352 vs 6 376 on 2 — what the shorter bubble
  // saves, the self-k-NN of the smaller source grids loses.)
  std::vector<std::pair<size_t, size_t>> chunks;
  for (size_t first = 0; first < n_pairs;) {
    size_t pts = 0, cnt = 0;
    while (first + cnt < n_pairs && (max_chunk_pairs <= 0 || (int)cnt < max_chunk_pairs) && cnt < 255) {
      const size_t add = pairs[first + cnt].n_source + pairs[first + cnt].n_target;
      if (cnt > 0 && pts + add > max_points) break;
      pts += add;
      cnt++;
    }
    chunks.push_back({first, cnt});
    first += cnt;
  }
  std::unique_ptr<ChunkIn> cur, nxt;
  if (!chunks.empty()) TRY(chunk_upload(c, pairs + chunks[0].first, (int)chunks[0].second, cur));
  for (size_t i = 0; i < chunks.size(); i++) {
    if (i + 1 < chunks.size()) TRY(chunk_upload(c, pairs + chunks[i + 1].first, (int)chunks[i + 1].second, nxt));
    TRY(batch_chunk(c, prm, pairs + chunks[i].first, *cur, want_fitness, fitness_max_range, lazy, out + chunks[i].first));
    cur = std::move(nxt);
  }
  return RGC_OK;
}

int rgc_batch_last_stage_ms(const rgc_ctx* c, float* ms6, int* rounds) {
  if (!c || !ms6) return RGC_ERR_INVALID;
  for (int i = 0; i < 6; i++) ms6[i] = g_batch_stats.ms[i];
  if (rounds) *rounds = g_batch_stats.rounds;
  return RGC_OK;
}

}  // extern "C"
