// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.hpp header).  PARITY UNPINNED.
//
// Exact k-nearest-neighbour search standing in for pcl::search::KdTree -> pcl::KdTreeFLANN ->
// flann::KDTreeSingleIndex<L2_Simple<float>> (un-vendored; call sites
// rgc_slam/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:133 and :254).
// Contract restated: exact kNN, ascending by squared distance computed in
// float as ((dx*dx)+dy*dy)+dz*dz (flann::L2_Simple accumulates dimension by dimension),
// k clamped to N, leaf size 15 (PCL's KDTreeSingleIndexParams(15)), points reordered.
// Tie policy (ours, FLANN's is traversal-order): (d2, index) lexicographic ascending.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace orc {

struct KdTree {
  struct Node {
    int left, right;      // children (inner) or [left,right) range into reordered points (leaf)
    int dim;              // -1 for leaf
    float divlow, divhigh;
  };
  std::vector<Node> nodes;
  std::vector<float> pts;   // reordered xyz (3 floats each)
  std::vector<int> ids;     // original index of reordered point
  int n = 0;
  static constexpr int kLeaf = 15;

  void build(const float* xyzw, int n_, int stride_floats = 4) {
    n = n_;
    ids.resize(n);
    for (int i = 0; i < n; i++) ids[i] = i;
    std::vector<float> src(3 * (size_t)n);
    for (int i = 0; i < n; i++)
      for (int d = 0; d < 3; d++) src[3 * (size_t)i + d] = xyzw[(size_t)i * stride_floats + d];
    nodes.clear();
    nodes.reserve(2 * (n / kLeaf + 2));
    if (n > 0) build_rec(src, 0, n);
    pts.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++)
      for (int d = 0; d < 3; d++) pts[3 * (size_t)i + d] = src[3 * (size_t)ids[i] + d];
  }

  int build_rec(const std::vector<float>& src, int lo, int hi) {
    int me = (int)nodes.size();
    nodes.push_back(Node());
    if (hi - lo <= kLeaf) {
      nodes[me] = {lo, hi, -1, 0.f, 0.f};
      return me;
    }
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; i++)
      for (int d = 0; d < 3; d++) {
        float v = src[3 * (size_t)ids[i] + d];
        mn[d] = std::min(mn[d], v);
        mx[d] = std::max(mx[d], v);
      }
    int dim = 0;
    for (int d = 1; d < 3; d++)
      if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    int mid = (lo + hi) / 2;
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int a, int b) {
      float va = src[3 * (size_t)a + dim], vb = src[3 * (size_t)b + dim];
      return va < vb || (va == vb && a < b);
    });
    float divlow = -1e30f, divhigh = 1e30f;
    for (int i = lo; i < mid; i++) divlow = std::max(divlow, src[3 * (size_t)ids[i] + dim]);
    for (int i = mid; i < hi; i++) divhigh = std::min(divhigh, src[3 * (size_t)ids[i] + dim]);
    int l = build_rec(src, lo, mid);
    int r = build_rec(src, mid, hi);
    nodes[me] = {l, r, dim, divlow, divhigh};
    return me;
  }

  struct Result {
    int k, count;
    float* d2;
    int* idx;
    inline bool full() const { return count == k; }
    inline float worst() const { return full() ? d2[k - 1] : 3.0e38f; }
    inline void insert(float d, int id) {
      if (full()) {
        if (!(d < d2[k - 1] || (d == d2[k - 1] && id < idx[k - 1]))) return;
      }
      int j = full() ? k - 1 : count++;
      while (j > 0 && (d2[j - 1] > d || (d2[j - 1] == d && idx[j - 1] > id))) {
        d2[j] = d2[j - 1];
        idx[j] = idx[j - 1];
        j--;
      }
      d2[j] = d;
      idx[j] = id;
    }
  };

  void search_rec(int ni, const float* q, float mindist, float* dists, Result& res) const {
    const Node& nd = nodes[ni];
    if (nd.dim < 0) {
      for (int i = nd.left; i < nd.right; i++) {
        const float* p = &pts[3 * (size_t)i];
        float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
        float d = dx * dx;
        d = d + dy * dy;
        d = d + dz * dz;
        res.insert(d, ids[i]);
      }
      return;
    }
    float val = q[nd.dim];
    float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int best, other;
    float cut;
    if (diff1 + diff2 < 0) {
      best = nd.left;
      other = nd.right;
      cut = diff2 * diff2;
    } else {
      best = nd.right;
      other = nd.left;
      cut = diff1 * diff1;
    }
    search_rec(best, q, mindist, dists, res);
    float dst = dists[nd.dim];
    float md = mindist + cut - dst;
    dists[nd.dim] = cut;
    // conservative prune (<= with slack) so that equal-distance / smaller-index candidates are never lost
    if (md * 0.999999f <= res.worst()) search_rec(other, q, md, dists, res);
    dists[nd.dim] = dst;
  }

  // returns number found (= min(k, n)); outputs ascending by (d2, idx)
  int knn(const float* q, int k, int* idx, float* d2) const {
    if (k > n) k = n;
    if (k <= 0) return 0;
    Result res{k, 0, d2, idx};
    float dists[3] = {0.f, 0.f, 0.f};
    search_rec(0, q, 0.f, dists, res);
    return res.count;
  }
};

}  // namespace orc
