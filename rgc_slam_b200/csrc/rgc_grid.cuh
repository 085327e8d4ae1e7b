// rgc_grid.cuh — Morton-ordered multi-level voxel hash and the exact kNN search that walks it.
//
// Replaces pcl::search::KdTree::nearestKSearch at the reference's two call sites
// (rgc_slam/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:133 1-NN per LM iteration,
//  :254 k=20 self-kNN for covariances).
//
// Data layout in HBM (one `GridView` per cloud):
//   pts[n]        float4, points sorted by the 3D Morton code of their finest-level cell;
//                 .w carries the ORIGINAL index (int bits) so a candidate is one 16-B load.
//   level l table open-addressing hash {key = morton >> 3l, start, end} (16-B slots): the
//                 points of any cell at any level are one contiguous [start,end) range of pts
//                 because Morton order nests (an octree laid out in an array).
// Cell size at level l is s0 * 2^l; level nbits is a single root cell.
//
// Search = for the query's cell c at level l, scan the 3x3x3 block around c, keep the best k by
// (d2, original index), and stop as soon as the k-th distance is provably smaller than the
// distance to the nearest block face that still has cells behind it; otherwise go one level
// coarser (8x volume).  The start level is found by climbing own-cell counts (1 lookup/level).
// Result is the exact kNN under the total order (d2, index) — independent of scan order, so
// no sort stability or atomics ordering can change it.
#pragma once
#include "rgc_common.cuh"

namespace rgc {

constexpr int kMaxLevels = 22;          // nbits <= 21  (3*21 = 63-bit Morton keys)
constexpr uint64_t kEmptyKey = ~0ull;

struct GridSlot {  // 16 bytes, loaded as one uint4
  uint64_t key;
  uint32_t start, end;
};

struct GridView {
  const F4* pts;   // sorted points, w = original index bits
  int n;
  float ox, oy, oz;   // origin
  float s0, inv_s0;   // finest cell size
  float margin;       // slack (metres) covering float rounding of the cell-coordinate map
  int nbits;          // cells per axis at level 0 = 1 << nbits
  int nlevels;        // nbits + 1
  const GridSlot* table[kMaxLevels];
  uint32_t mask[kMaxLevels];  // table size - 1 (power of two)
};

// Grid geometry from a bounding box (host side; shared by the library and tests/hostsim so both
// place every point in the same cell).  Sets origin, cell size, nbits/nlevels and the margin.
inline void grid_geometry(const float mn[3], const float mx[3], float cell, GridView& v) {
  float s0 = cell > 0.f ? cell : 0.2f;
  float extent = 0.f, maxabs = 0.f;
  for (int a = 0; a < 3; a++) {
    extent = fmaxf(extent, mx[a] - mn[a]);
    maxabs = fmaxf(maxabs, fmaxf(fabsf(mn[a]), fabsf(mx[a])));
  }
  extent += 2.f * s0;
  if (extent / s0 > (float)(1 << 21)) s0 = extent / (float)(1 << 21) * 1.001f;
  int nbits = 1;
  while ((float)(1 << nbits) * s0 < extent && nbits < 21) nbits++;
  v.ox = mn[0] - s0;
  v.oy = mn[1] - s0;
  v.oz = mn[2] - s0;
  v.s0 = s0;
  v.inv_s0 = 1.0f / s0;
  v.margin = 2e-6f * (maxabs + extent) + 1e-6f;
  v.nbits = nbits;
  v.nlevels = nbits + 1;
}

// ---- Morton ------------------------------------------------------------------------------------
RGC_HD uint64_t spread3(uint32_t v) {  // 21 bits -> every third bit
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
RGC_HD uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }

RGC_HD uint64_t mix64(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// fine-level integer cell coordinate of a float coordinate.  Monotone non-decreasing in x
// (float subtract, multiply and floor are monotone), which is what the face bounds rely on.
RGC_HD int cell_coord(float x, float o, float inv_s) {
  float t = fmul(fsub(x, o), inv_s);
  t = t < -1.0e9f ? -1.0e9f : (t > 1.0e9f ? 1.0e9f : t);
  return (int)floorf(t);
}

RGC_HD GridSlot load_slot(const GridSlot* p) {
#if defined(__CUDA_ARCH__)
  uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  GridSlot s;
  s.key = (uint64_t)v.x | ((uint64_t)v.y << 32);
  s.start = v.z;
  s.end = v.w;
  return s;
#else
  return *p;
#endif
}

RGC_HD F4 load_pt(const F4* p) {
#if defined(__CUDA_ARCH__)
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return F4{v.x, v.y, v.z, v.w};
#else
  return *p;
#endif
}

// cell (cx,cy,cz) at level l -> [start,end) ; false if the cell is empty / out of range
RGC_HD bool grid_lookup(const GridView& g, int l, int cx, int cy, int cz, uint32_t& start, uint32_t& end) {
  const int ncell = 1 << (g.nbits - l);
  if ((unsigned)cx >= (unsigned)ncell || (unsigned)cy >= (unsigned)ncell || (unsigned)cz >= (unsigned)ncell) return false;
  const uint64_t key = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
  const GridSlot* tab = g.table[l];
  const uint32_t mask = g.mask[l];
  uint32_t h = (uint32_t)mix64(key) & mask;
  for (;;) {
    GridSlot s = load_slot(tab + h);
    if (s.key == key) {
      start = s.start;
      end = s.end;
      return true;
    }
    if (s.key == kEmptyKey) return false;
    h = (h + 1) & mask;
  }
}

// ---- bounded sorted list of the k best candidates, kept in registers ---------------------------
// KCAP slots, ascending by (d2, original index).  The k live entries occupy the LAST k slots, so
// the current k-th best always sits in the static slot KCAP-1 (a run-time position would make
// the compiler index the arrays dynamically and demote them to local memory); the leading
// KCAP-k slots are -inf sentinels that no candidate can precede.  id[] = sorted position in
// g.pts (-1 = empty, -2 = sentinel).  Exact distance ties are ordered by the ORIGINAL index,
// fetched lazily from pts[].w only when two distances compare equal.
template <int KCAP>
struct TopK {
  float d[KCAP];
  int id[KCAP];
  float lim;  // external pruning radius (inclusive): candidates with d2 > lim are never needed
  int k;

  RGC_HD void reset(int k_, float lim_) {
    k = k_;
    lim = lim_;
#pragma unroll
    for (int j = 0; j < KCAP; j++) {
      const bool live = j >= KCAP - k_;
      d[j] = live ? INFINITY : -INFINITY;
      id[j] = live ? -1 : -2;
    }
  }
  RGC_HD bool full() const { return id[KCAP - 1] >= 0; }
  RGC_HD float worst() const { return d[KCAP - 1]; }
  // (dn, orig_n) strictly before (dj, orig(idj)) ?
  RGC_HD bool before(float dn, int orig_n, float dj, int idj, const F4* pts) const {
    if (dn < dj) return true;
    if (dn > dj) return false;
    if (idj < 0) return idj == -1;  // unreachable for finite dn: empty slots hold +inf, sentinels -inf
    return orig_n < f2i_bits(load_pt(pts + idj).w);
  }
  RGC_HD void insert(float dn, int pos, int orig, const F4* pts) {
    if (dn > lim) return;
    // new entry must precede the current k-th best (static slot KCAP-1)
    bool b_j = before(dn, orig, d[KCAP - 1], id[KCAP - 1], pts);
    if (!b_j) return;
    // walk from the tail: slot j takes slot j-1 if the new entry precedes j-1, else takes the
    // new entry if it precedes j.
#pragma unroll
    for (int j = KCAP - 1; j >= 1; j--) {
      const bool b_jm1 = before(dn, orig, d[j - 1], id[j - 1], pts);
      if (b_jm1) {
        d[j] = d[j - 1];
        id[j] = id[j - 1];
      } else if (b_j) {
        d[j] = dn;
        id[j] = pos;
      }
      b_j = b_jm1;
    }
    if (b_j) {
      d[0] = dn;
      id[0] = pos;
    }
  }
  // j-th best (0-based) lives in slot KCAP-k+j; callers read it with a static unrolled loop
};

struct SearchStats {
  int levels, lookups, candidates;
};

// Exact kNN of (qx,qy,qz) in grid g.  `max_d2`: candidates with d2 > max_d2 are never needed
// (pass +inf for an unbounded search).  `level_hint` (>=0) starts the climb at that level.
// Returns the level the search finished at (useful as the next hint).
template <int KCAP>
RGC_HD int knn_search(const GridView& g, float qx, float qy, float qz, int k, float max_d2, int level_hint, TopK<KCAP>& top,
                      SearchStats* st = nullptr) {
  const int fx = cell_coord(qx, g.ox, g.inv_s0);
  const int fy = cell_coord(qy, g.oy, g.inv_s0);
  const int fz = cell_coord(qz, g.oz, g.inv_s0);
  const int top_level = g.nlevels - 1;

  // ---- pick the start level: the parent cell at level l+1 lies inside the 3x3x3 block of
  // level l, so the first level whose own-cell count reaches k gives a block that fills the list.
  int l = level_hint < 0 ? 0 : (level_hint > top_level ? top_level : level_hint);
  if (level_hint < 0) {
    int lv = 1;
    for (; lv <= top_level; lv++) {
      uint32_t s, e;
      if (st) st->lookups++;
      if (grid_lookup(g, lv, fx >> lv, fy >> lv, fz >> lv, s, e) && (int)(e - s) >= k) break;
    }
    l = lv - 1;
    if (l > top_level) l = top_level;
  }

  float prune = max_d2;
  for (;; l++) {
    if (l > top_level) l = top_level;
    const int cx = fx >> l, cy = fy >> l, cz = fz >> l;
    top.reset(k, prune);  // prune carries the best known k-th distance (or the caller's radius)
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          uint32_t s, e;
          if (st) st->lookups++;
          if (!grid_lookup(g, l, cx + dx, cy + dy, cz + dz, s, e)) continue;
          for (uint32_t p = s; p < e; p++) {
            F4 c = load_pt(g.pts + p);
            float d2 = dist2_ref(qx, qy, qz, c.x, c.y, c.z);
            if (st) st->candidates++;
            top.insert(d2, (int)p, f2i_bits(c.w), g.pts);
          }
        }
    if (st) st->levels++;
    // ---- termination: distance to the nearest face of the block that still has cells behind it
    const int ncell = 1 << (g.nbits - l);
    const float cs = g.s0 * (float)(1 << l);
    float gap = INFINITY;
    {
      const int c3[3] = {cx, cy, cz};
      const float q3[3] = {qx, qy, qz};
      const float o3[3] = {g.ox, g.oy, g.oz};
#pragma unroll
      for (int a = 0; a < 3; a++) {
        if (c3[a] - 1 > 0) {  // cells exist below the low face
          float lo = o3[a] + (float)(c3[a] - 1) * cs;
          float ga = q3[a] - lo - g.margin;
          gap = ga < gap ? ga : gap;
        }
        if (c3[a] + 2 < ncell) {  // cells exist above the high face
          float hi = o3[a] + (float)(c3[a] + 2) * cs;
          float ga = hi - q3[a] - g.margin;
          gap = ga < gap ? ga : gap;
        }
      }
    }
    if (gap == INFINITY) break;  // block covers the whole grid
    if (gap > 0.f) {
      const float gap2 = gap * gap * 0.999999f;
      const bool full = top.full();
      if (full && top.worst() < gap2) break;  // k-th best is closer than anything outside
      if (gap2 > max_d2) break;               // everything the caller can accept has been seen
      if (full) prune = top.worst();
    }
    if (l == top_level) break;  // unreachable: the top block always covers the grid
  }
  return l;
}

}  // namespace rgc
