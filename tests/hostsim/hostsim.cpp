// tests/hostsim/hostsim.cpp — TEST-ONLY CPU simulation of the kernels' per-thread logic.
//
// Compiles the host/device headers of the product (rgc_grid.cuh, rgc_math.cuh) with g++ and
// runs one "thread" per point in plain loops, building the Morton order and level tables the way
// the CUDA build pipeline does.  Lets the CPU test-suite check the grid search, covariance and
// linearize arithmetic against the oracle in a container that has no GPU.  It is never linked
// into librgc_gicp.so and is not a fallback: the product library has no CPU path.
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <vector>

#include "../../rgc_slam_b200/csrc/rgc_grid.cuh"
#include "../../rgc_slam_b200/csrc/rgc_math.cuh"
#include "../../rgc_slam_b200/csrc/rgc_mapping.cuh"
#include "../../rgc_slam_b200/csrc/rgc_preprocess.cuh"

using namespace rgc;

struct SimCloud {
  std::vector<F4> sorted;
  std::vector<std::vector<GridSlot>> tables;
  GridView v{};
};

static void sim_build(const float* xyzw, int n, float cell, SimCloud& c) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], xyzw[4 * (size_t)i + a]);
      mx[a] = std::max(mx[a], xyzw[4 * (size_t)i + a]);
    }
  GridView& v = c.v;
  v.n = n;
  grid_geometry(mn, mx, cell, v);
  const int hi = (1 << v.nbits) - 1;
  std::vector<uint64_t> keys(n);
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) {
    int cx = std::min(std::max(cell_coord(xyzw[4 * (size_t)i], v.inv_s0, v.bias), 0), hi);
    int cy = std::min(std::max(cell_coord(xyzw[4 * (size_t)i + 1], v.inv_s0, v.bias), 0), hi);
    int cz = std::min(std::max(cell_coord(xyzw[4 * (size_t)i + 2], v.inv_s0, v.bias), 0), hi);
    keys[i] = morton3(cx, cy, cz);
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
  c.sorted.resize(n);
  std::vector<uint64_t> ks(n);
  for (int i = 0; i < n; i++) {
    int o = order[i];
    c.sorted[i] = F4{xyzw[4 * (size_t)o], xyzw[4 * (size_t)o + 1], xyzw[4 * (size_t)o + 2], i2f_bits(o)};
    ks[i] = keys[o];
  }
  c.tables.assign(v.nlevels, {});
  for (int l = 0; l < v.nlevels; l++) {
    size_t cells = 0;
    for (int i = 0; i < n; i++)
      if (i == 0 || (ks[i] >> (3 * l)) != (ks[i - 1] >> (3 * l))) cells++;
    size_t s = 8;
    while (s < 2 * cells) s <<= 1;
    auto& t = c.tables[l];
    t.assign(s, GridSlot{kEmptyKey, 0, 0});
    uint32_t mask = (uint32_t)(s - 1);
    int lg = 0;
    while (((size_t)1 << lg) < s) lg++;
    v.shift[l] = (uint32_t)(64 - lg);
    int start = 0;
    for (int i = 1; i <= n; i++) {
      if (i == n || (ks[i] >> (3 * l)) != (ks[i - 1] >> (3 * l))) {
        uint64_t key = ks[i - 1] >> (3 * l);
        uint32_t h = slot_of(key, v.shift[l]);
        while (t[h].key != kEmptyKey) h = (h + 1) & mask;
        t[h] = GridSlot{key, (uint32_t)start, (uint32_t)i};
        start = i;
      }
    }
    v.table[l] = t.data();
    v.mask[l] = mask;
  }
  // occupied-children masks (k_child_masks)
  for (int l = 0; l + 1 < v.nlevels; l++)
    for (int i = 0; i < n; i++)
      if (i == 0 || (ks[i] >> (3 * l)) != (ks[i - 1] >> (3 * l))) {
        uint64_t ck = ks[i] >> (3 * l), pk = ck >> 3;
        auto& t = c.tables[l + 1];
        uint32_t mask = (uint32_t)(t.size() - 1);
        uint32_t hh = slot_of(pk, v.shift[l + 1]);
        while ((t[hh].key & kKeyMask) != pk) hh = (hh + 1) & mask;
        t[hh].key |= (uint64_t)1 << (56 + (int)(ck & 7));
      }
  v.pts = c.sorted.data();
  v.inv = nullptr;
}

static void sim_knn_t(const SimCloud& c, const float* q, int m, int k, int* idx, float* d2, long long* stats) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int t = 0; t < m; t++) {
    float hd[64];
    int hi[64];
    HeapK top;
    top.init(hd, hi, 1);
    SearchStats st{0, 0, 0};
    knn_search(c.v, q[4 * (size_t)t], q[4 * (size_t)t + 1], q[4 * (size_t)t + 2], k, INFINITY, -1, top, &st);
    top.sort_ascending(c.v.pts);
    for (int j = 0; j < k; j++) {
      const bool have = j < top.cnt;
      idx[(size_t)t * k + j] = have ? f2i_bits(c.sorted[hi[j]].w) : -1;
      d2[(size_t)t * k + j] = have ? hd[j] : INFINITY;
    }
    if (stats) {
#pragma omp atomic
      stats[0] += st.nodes;
#pragma omp atomic
      stats[1] += st.lookups;
#pragma omp atomic
      stats[2] += st.candidates;
    }
  }
}

extern "C" {

int sim_knn(const float* pts, int n, const float* queries, int m, int k, int* idx, float* d2, float cell, long long* stats) {
  SimCloud c;
  sim_build(pts, n, cell, c);
  if (stats) stats[0] = stats[1] = stats[2] = 0;
  if (k > 64) return -1;
  sim_knn_t(c, queries, m, k, idx, d2, stats);
  return c.v.nlevels;
}

// covariances (column-major == row-major 4x4, symmetric) in ORIGINAL order, from given kNN lists
void sim_covs(const float* pts, int n, const int* knn_idx, int k, int method, double* covs16) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    int found = 0;
    while (found < k && knn_idx[(size_t)i * k + found] >= 0) found++;
    Sym3 c = covariance_from_points(found, k, [&](int j) {
      const float* p = &pts[4 * (size_t)knn_idx[(size_t)i * k + j]];
      return F4{p[0], p[1], p[2], p[3]};
    });
    Sym3 r = regularize_cov(c, method);
    double* o = &covs16[16 * (size_t)i];
    for (int j = 0; j < 16; j++) o[j] = 0.0;
    o[0] = r.xx; o[1] = r.xy; o[2] = r.xz;
    o[4] = r.xy; o[5] = r.yy; o[6] = r.yz;
    o[8] = r.xz; o[9] = r.yz; o[10] = r.zz;
  }
}

// one linearize: T row-major 4x4 double; covs as 16-double matrices in original order.
// out: err, H (row-major 36), b (6), corr (original target index or -1)
void sim_linearize(const float* src, int ns, const float* tgt, int nt, const double* covA16, const double* covB16, const double* T, float thr,
                   float cell, double* err, double* H, double* b, int* corr) {
  SimCloud c;
  sim_build(tgt, nt, cell, c);
  Rt Td;
  float Tf[12];
  for (int i = 0; i < 12; i++) {
    Td.m[i] = T[i];
    Tf[i] = (float)T[i];
  }
  const float thr2 = thr * thr;
  double acc[kAccN] = {0};
  for (int i = 0; i < ns; i++) {
    const float* p = &src[4 * (size_t)i];
    float qx, qy, qz;
    transform_f(Tf, p[0], p[1], p[2], qx, qy, qz);
    Best1 top;
    knn_search(c.v, qx, qy, qz, 1, thr2, -1, top);
    int pos = (top.id0 >= 0 && top.d0 < thr2) ? top.id0 : -1;
    corr[i] = pos >= 0 ? f2i_bits(c.sorted[pos].w) : -1;
    if (pos < 0) continue;
    const double* a = &covA16[16 * (size_t)i];
    const double* bb = &covB16[16 * (size_t)corr[i]];
    Sym3 CA{a[0], a[1], a[2], a[5], a[6], a[10]}, CB{bb[0], bb[1], bb[2], bb[5], bb[6], bb[10]};
    Sym3 M = gicp_mahalanobis(Td, CA, CB);
    F4 q = c.sorted[pos];
    gicp_point_terms(Td, M, p[0], p[1], p[2], q.x, q.y, q.z, acc);
  }
  *err = acc[0];
  int o = 1;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) {
      H[i * 6 + j] = H[j * 6 + i] = acc[o];
      o++;
    }
  for (int i = 0; i < 6; i++) b[i] = acc[22 + i];
}

}  // extern "C"

// host-side LM helpers of the product (rgc_lm.hpp) exposed for the CPU tests
#include "../../rgc_slam_b200/csrc/rgc_lm.hpp"
extern "C" {
void sim_solve_ldlt6(const double* A, const double* rhs, double* x) { rgc::lm::solve_ldlt6(A, rhs, x); }
void sim_se3_delta(const double* d, double* delta16) { rgc::lm::se3_delta(d, delta16); }
int sim_is_converged(const double* delta16, double rot_eps, double trans_eps) { return rgc::lm::is_converged(delta16, rot_eps, trans_eps) ? 1 : 0; }
}

// ---- pre-step and mapping association: the product's host/device arithmetic, one "thread" per point ----
extern "C" {

// rgc_gicp.cu pre_filter + k_pre_ingest: de-skew of every point (q = q_last_curr as w, x, y, z)
void sim_deskew(const float* xyzi, int n, const double* q, const double* t3, float period, float* out) {
  DeskewParams D{};
  const double n2 = ((q[1] * q[1] + q[2] * q[2]) + q[3] * q[3]) + q[0] * q[0];
  D.enabled = 1;
  if (n2 > 0.0) {
    D.iw = q[0] / n2;
    D.ix = -q[1] / n2;
    D.iy = -q[2] / n2;
    D.iz = -q[3] / n2;
  }
  D.tx = t3[0];
  D.ty = t3[1];
  D.tz = t3[2];
  D.scan_period = period;
  for (int i = 0; i < n; i++) {
    float x = xyzi[4 * (size_t)i], y = xyzi[4 * (size_t)i + 1], z = xyzi[4 * (size_t)i + 2];
    const float inten = xyzi[4 * (size_t)i + 3];
    deskew_point(D, inten, x, y, z);
    out[4 * (size_t)i] = x;
    out[4 * (size_t)i + 1] = y;
    out[4 * (size_t)i + 2] = z;
    out[4 * (size_t)i + 3] = inten;
  }
}

// pre_filter's grid set-up + vg_index + the stable sort / in-order float sums of k_vg_centroid
int sim_voxel_grid(const float* xyzi, int n, float leaf, float* out) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], xyzi[4 * (size_t)i + a]);
      mx[a] = std::max(mx[a], xyzi[4 * (size_t)i + a]);
    }
  VgGeom g;
  g.inv_leaf = 1.0f / leaf;
  long long dd[3];
  for (int a = 0; a < 3; a++) dd[a] = (long long)((mx[a] - mn[a]) * g.inv_leaf) + 1;
  if (dd[0] * dd[1] * dd[2] > 2147483647ll) {
    std::copy(xyzi, xyzi + 4 * (size_t)n, out);
    return n;
  }
  int div_b[3];
  for (int a = 0; a < 3; a++) {
    g.min_b[a] = (int)std::floor(mn[a] * g.inv_leaf);
    div_b[a] = (int)std::floor(mx[a] * g.inv_leaf) - g.min_b[a] + 1;
  }
  g.mul1 = div_b[0];
  g.mul2 = div_b[0] * div_b[1];
  std::vector<unsigned> key(n);
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) {
    key[i] = vg_index(g, xyzi[4 * (size_t)i], xyzi[4 * (size_t)i + 1], xyzi[4 * (size_t)i + 2]);
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  int m = 0;
  for (int s = 0; s < n;) {
    int e = s + 1;
    while (e < n && key[order[e]] == key[order[s]]) e++;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int j = s; j < e; j++) {
      const float* p = xyzi + 4 * (size_t)order[j];
      sx = fadd(sx, p[0]);
      sy = fadd(sy, p[1]);
      sz = fadd(sz, p[2]);
      si = fadd(si, p[3]);
    }
    const float cnt = (float)(e - s);
    out[4 * (size_t)m] = sx / cnt;
    out[4 * (size_t)m + 1] = sy / cnt;
    out[4 * (size_t)m + 2] = sz / cnt;
    out[4 * (size_t)m + 3] = si / cnt;
    m++;
    s = e;
  }
  return m;
}

// k_map_assoc on the host: pose transform, exact 5-NN through knn_search, edge / plane fit
void sim_map_assoc(const float* map_xyz1, int nm, const float* feats, int n, const double* q, const double* t3, int plane, int* valid, double* o1, double* o2) {
  SimCloud c;
  sim_build(map_xyz1, nm, 0.f, c);
  const PoseQ T{q[0], q[1], q[2], q[3], t3[0], t3[1], t3[2]};
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; i++) {
    valid[i] = 0;
    float qx, qy, qz;
    associate_to_map(T, feats[4 * (size_t)i], feats[4 * (size_t)i + 1], feats[4 * (size_t)i + 2], qx, qy, qz);
    float hd[8];
    int hi[8];
    HeapK top;
    top.init(hd, hi, 1);
    knn_search(c.v, qx, qy, qz, 5, INFINITY, -1, top);
    top.sort_ascending(c.v.pts);
    if (!(top.cnt == 5 && hd[4] < (plane ? 2.0f : 1.0f))) continue;
    double P[5][3];
    for (int j = 0; j < 5; j++) {
      const F4 p = c.sorted[hi[j]];
      P[j][0] = (double)p.x;
      P[j][1] = (double)p.y;
      P[j][2] = (double)p.z;
    }
    if (!plane) {
      double pa[3], pb[3];
      if (!edge_fit(P, pa, pb)) continue;
      valid[i] = 1;
      for (int r = 0; r < 3; r++) {
        o1[3 * (size_t)i + r] = pa[r];
        o2[3 * (size_t)i + r] = pb[r];
      }
    } else {
      double nrm[3], d;
      if (!plane_fit(P, nrm, d)) continue;
      valid[i] = 1;
      for (int r = 0; r < 3; r++) o1[3 * (size_t)i + r] = nrm[r];
      o2[i] = d;
    }
  }
}

}  // extern "C"

// HeapK64 (the tile kernel's per-lane container) driven on the host: k smallest of n keys, ascending
extern "C" void sim_heap64_topk(const unsigned long long* keys, int n, int k, int stride, unsigned long long* out) {
  std::vector<unsigned long long> store((size_t)k * stride, 0ull);
  HeapK64 hp;
  hp.init(store.data(), stride, k);
  for (int i = 0; i < n; i++) hp.insert(keys[i]);
  hp.sort_ascending();
  for (int j = 0; j < hp.cnt; j++) out[j] = store[(size_t)j * stride];
  for (int j = hp.cnt; j < k; j++) out[j] = ~0ull;
}
