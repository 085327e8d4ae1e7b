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
// Each slot also carries an 8-bit mask of its occupied children (top byte of the key word).
//
// Search (exact, density-adaptive):
//   1. bound  = largest d2 to k points that are adjacent in Morton order to the query (k loads) —
//               a valid upper bound on the k-th nearest distance;
//   2. roots  = the <= 3x3x3 cells, at the finest level whose cell edge >= sqrt(bound), that
//               intersect the ball(q, sqrt(bound));
//   3. descend each root depth-first, nearest child first, pruning every cell whose box is
//               farther than the current k-th best, scanning the contiguous point range of a cell
//               once it holds <= kLeafPoints points (or is at level 0).
// Work is proportional to the points near the ball, whatever the local density (a sparse query
// next to a dense region never scans the dense region).  The result is the exact kNN under the
// total order (d2, original index) — independent of scan order, sort stability or atomics.
#pragma once
#include "rgc_common.cuh"

namespace rgc {

constexpr int kMaxBits = 18;            // cells per axis <= 2^18 (54-bit Morton keys + 8-bit child mask)
constexpr int kMaxLevels = kMaxBits + 1;
constexpr uint64_t kKeyMask = (1ull << 56) - 1;
constexpr int kLeafPoints = 12;         // scan a cell directly once it holds this few points
constexpr int kStackCap = 40;           // DFS stack entries per thread (overflow => scan the cell)
constexpr uint64_t kEmptyKey = ~0ull;

struct GridSlot {  // 16 bytes, loaded as one uint4
  uint64_t key;  // [55:0] Morton key at this level, [63:56] occupied-children mask
  uint32_t start, end;
};

struct GridView {
  const int* inv;  // original index -> sorted position
  const F4* pts;   // sorted points, w = original index bits
  int n;
  float ox, oy, oz;   // origin (metres) = -bias * s0 on every axis: the cells are ABSOLUTE, see grid_geometry
  int bias;           // fine cell of coordinate x = floor(x * inv_s0) + bias
  float s0, inv_s0;   // finest cell size
  float margin;       // slack (metres) covering float rounding of the cell-coordinate map
  int nbits;          // cells per axis at level 0 = 1 << nbits
  int nlevels;        // nbits + 1
  const GridSlot* table[kMaxLevels];
  uint32_t mask[kMaxLevels];   // table size - 1 (power of two)
  uint32_t shift[kMaxLevels];  // 64 - log2(table size)
};

// Grid geometry from a bounding box (host side; shared by the library and tests/hostsim so both
// place every point in the same cell).  The grid is ABSOLUTE: the fine cell of a coordinate x is
// floor(x * inv_s0) + bias with bias = 2^(nbits-1), whatever the cloud — the bounding box only decides how
// many bits are needed.  Consequence: the Morton order of a cloud's points does not depend on what else
// shares the grid nor on nbits (a larger nbits only inserts, below each axis' top bit, bits that repeat
// its complement and never decide a comparison), so a cloud is sorted the same way alone and inside the
// multi-cloud grid of a batch, and every fixed-order reduction over its points adds in the same order.
// view fields of a grid with `nbits` bits per axis and finest cell s0
inline void grid_set(GridView& v, int nbits, float s0) {
  v.bias = 1 << (nbits - 1);
  v.ox = v.oy = v.oz = -(float)v.bias * s0;
  v.s0 = s0;
  v.inv_s0 = 1.0f / s0;
  v.nbits = nbits;
  v.nlevels = nbits + 1;
}
// slack (metres) covering the float rounding of the cell-coordinate map over this bounding box
inline float grid_margin(const float mn[3], const float mx[3]) {
  float extent = 0.f, maxabs = 0.f;
  for (int a = 0; a < 3; a++) {
    extent = fmaxf(extent, mx[a] - mn[a]);
    maxabs = fmaxf(maxabs, fmaxf(fabsf(mn[a]), fabsf(mx[a])));
  }
  return 2e-6f * (2.f * maxabs + extent) + 1e-6f;
}
// smallest (nbits, s0) that holds the bounding box
inline void grid_bits(const float mn[3], const float mx[3], float cell, int max_bits, int& nbits_out, float& s0_out) {
  float s0 = cell > 0.f ? cell : 0.05f;
  float maxabs = 0.f;
  for (int a = 0; a < 3; a++) maxabs = fmaxf(maxabs, fmaxf(fabsf(mn[a]), fabsf(mx[a])));
  // cells floor(x / s0) of all points must lie in [-bias, bias): bias * s0 > maxabs + s0
  int nbits = 2;
  for (;;) {
    while (nbits < max_bits && (float)(1 << (nbits - 1)) * s0 <= maxabs + 2.f * s0) nbits++;
    if ((float)(1 << (nbits - 1)) * s0 > maxabs + 2.f * s0) break;
    s0 *= 2.f;  // coarser cells (a power-of-two ladder) once max_bits cannot cover the extent
  }
  nbits_out = nbits;
  s0_out = s0;
}
inline void grid_geometry(const float mn[3], const float mx[3], float cell, GridView& v, int max_bits = kMaxBits) {
  int nbits;
  float s0;
  grid_bits(mn, mx, cell, max_bits, nbits, s0);
  grid_set(v, nbits, s0);
  v.margin = grid_margin(mn, mx);
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

// table slot of a Morton key: Fibonacci (multiplicative) hashing — one 64-bit multiply, top bits.
// `shift` = 64 - log2(table size).  (murmur's fmix64, used first, cost two multiplies and three
// xor-shifts on the critical path of every probe.)
RGC_HD uint32_t slot_of(uint64_t key, uint32_t shift) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> shift); }

// 2^l as a float, built from its bit pattern (l in [-126, 127]): the search code needs cell edges s0 * 2^l at
// every node, and an int -> float conversion is a quarter-rate instruction on the critical path
RGC_HD float pow2f(int l) { return i2f_bits((127 + l) << 23); }

// smallest level l in [0, top_level] whose cell edge s0 * 2^l is >= r (top_level if none; r = +inf included).
// s0 * 2^l is exact, so with l = exponent(r) - exponent(s0) the product has r's exponent and is >= r iff
// s0's mantissa is: one comparison instead of a loop over the levels (which was 14 % of the hinted
// correspondence search's instructions, ncu source page of round 2).
RGC_HD int root_level(float s0, float r, int top_level) {
  if (!(r > s0)) return 0;
  if (!(r < INFINITY)) return top_level;
  int l = ((f2i_bits(r) >> 23) & 0xff) - ((f2i_bits(s0) >> 23) & 0xff);
  if (l > top_level) return top_level;
  if (!(s0 * pow2f(l) >= r)) l++;
  return l < top_level ? l : top_level;
}

// fine-level integer cell coordinate of a float coordinate: floor(x * inv_s) + bias.  Monotone
// non-decreasing in x (float multiply and floor are monotone), which is what the face bounds rely on;
// the only rounding is that of x * inv_s (<= 6e-8 |x| metres, covered by GridView::margin).
RGC_HD int cell_coord(float x, float inv_s, int bias) {
  float t = fmul(x, inv_s);
  t = t < -1.0e9f ? -1.0e9f : (t > 1.0e9f ? 1.0e9f : t);
  return (int)floorf(t) + bias;
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

// A grid may hold SEVERAL clouds (batched registration, rgc_batch.cuh): the clouds share the geometry,
// cloud c's points are the sorted positions [lo, hi) and its cells carry the key prefix
// c << 3 * nbits (at level l: prefix >> 3 l), so a search confined to one cloud never sees another
// cloud's points.  A single-cloud grid is the range {0, n, 0}.
struct CloudRange {
  int lo, hi;
  uint64_t prefix;
};
RGC_HD uint64_t prefix_at(const CloudRange& cr, int l) { return cr.prefix >> (3 * l); }

// cell (cx,cy,cz) at level l -> [start,end) + child mask ; false if the cell is empty / out of range
RGC_HD bool grid_lookup(const GridView& g, int l, int cx, int cy, int cz, uint32_t& start, uint32_t& end, uint32_t& cmask, uint64_t pfx = 0ull) {
  const int ncell = 1 << (g.nbits - l);
  if ((unsigned)cx >= (unsigned)ncell || (unsigned)cy >= (unsigned)ncell || (unsigned)cz >= (unsigned)ncell) return false;
  const uint64_t key = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz) | pfx;
  const GridSlot* tab = g.table[l];
  const uint32_t mask = g.mask[l];
  uint32_t h = slot_of(key, g.shift[l]);
  for (;;) {
    GridSlot s = load_slot(tab + h);
    if (s.key == kEmptyKey) return false;
    if ((s.key & kKeyMask) == key) {
      start = s.start;
      end = s.end;
      cmask = (uint32_t)(s.key >> 56);
      return true;
    }
    h = (h + 1) & mask;
  }
}

// same, for a cell known by its Morton key (children of a visited cell: key = parent << 3 | octant)
RGC_HD bool grid_lookup_key(const GridView& g, int l, uint64_t key, uint32_t& start, uint32_t& end, uint32_t& cmask) {
  const GridSlot* tab = g.table[l];
  const uint32_t mask = g.mask[l];
  uint32_t h = slot_of(key, g.shift[l]);
  for (;;) {
    GridSlot s = load_slot(tab + h);
    if (s.key == kEmptyKey) return false;
    if ((s.key & kKeyMask) == key) {
      start = s.start;
      end = s.end;
      cmask = (uint32_t)(s.key >> 56);
      return true;
    }
    h = (h + 1) & mask;
  }
}

// conservative squared distance from q to the box of cell (cx,cy,cz) at level l: never larger
// than the reference-arithmetic d2 of any point stored in that cell
RGC_HD float box_dist2(const GridView& g, int l, int cx, int cy, int cz, float qx, float qy, float qz) {
  const float cs = g.s0 * pow2f(l);
  const float lox = g.ox + (float)cx * cs, loy = g.oy + (float)cy * cs, loz = g.oz + (float)cz * cs;
  float gx = fmaxf(fmaxf(lox - qx, qx - (lox + cs)) - g.margin, 0.f);
  float gy = fmaxf(fmaxf(loy - qy, qy - (loy + cs)) - g.margin, 0.f);
  float gz = fmaxf(fmaxf(loz - qz, qz - (loz + cs)) - g.margin, 0.f);
  return (gx * gx + gy * gy + gz * gz) * 0.999999f;
}

// ---- the k best candidates so far --------------------------------------------------------------
// Binary max-heap ordered by (d2, original index): the root is the current k-th best, so the
// pruning bound is O(1) and an insertion is one sift (<= log2 k steps) instead of a k-slot shift.
// Storage is caller-provided and strided (device: one column of a shared-memory array per thread,
// dynamic indexing is free there; an ncu profile of the register-resident sorted-list version
// showed 62 % of all instructions in its 20-slot insertion at 4-5 active lanes).
// id[] = sorted position in g.pts.  Exact distance ties are ordered by the ORIGINAL index, fetched
// lazily from pts[].w only when two distances compare equal.
struct HeapK {
  float* d;
  int* id;
  int stride;
  int k, cnt;
  float lim;  // external pruning radius (inclusive): candidates with d2 > lim are never needed

  RGC_HD void init(float* d_, int* id_, int stride_) {
    d = d_;
    id = id_;
    stride = stride_;
  }
  RGC_HD void reset(int k_, float lim_) {
    k = k_;
    cnt = 0;
    lim = lim_;
  }
  RGC_HD bool full() const { return cnt == k; }
  RGC_HD float worst() const { return d[0]; }  // valid when full()
  RGC_HD float bound() const { return cnt == k ? fminf(lim, d[0]) : lim; }
  // (da, a) ordered after (db, b)?   a/b are sorted positions; orig_a may be supplied (>= 0)
  RGC_HD static bool after(float da, int orig_a, float db, int pb, const F4* pts) {
    if (da > db) return true;
    if (da < db) return false;
    return orig_a > f2i_bits(load_pt(pts + pb).w);
  }
  RGC_HD void insert(float dn, int pos, int orig, const F4* pts) {
    if (dn > lim) return;
    if (cnt < k) {  // sift up
      int j = cnt++;
      while (j > 0) {
        const int p = (j - 1) >> 1;
        const float dp = d[p * stride];
        const int ip = id[p * stride];
        if (!after(dn, orig, dp, ip, pts)) break;
        d[j * stride] = dp;
        id[j * stride] = ip;
        j = p;
      }
      d[j * stride] = dn;
      id[j * stride] = pos;
      return;
    }
    // full: the new entry must precede the current worst (the root), then sifts down from it
    {
      const float dr = d[0];
      if (dn > dr) return;
      if (dn == dr && !(orig < f2i_bits(load_pt(pts + id[0]).w))) return;
    }
    sift_down_from_root(dn, pos, orig, k, pts);
  }
  // same as insert(), but the original index of the new entry is fetched only if a tie needs it
  RGC_HD void insert_lazy(float dn, int pos, const F4* pts) {
    if (dn > lim) return;
    if (cnt < k) {
      int j = cnt++;
      while (j > 0) {
        const int p = (j - 1) >> 1;
        const float dp = d[p * stride];
        const int ip = id[p * stride];
        bool up = dn > dp;
        if (dn == dp) up = f2i_bits(load_pt(pts + pos).w) > f2i_bits(load_pt(pts + ip).w);
        if (!up) break;
        d[j * stride] = dp;
        id[j * stride] = ip;
        j = p;
      }
      d[j * stride] = dn;
      id[j * stride] = pos;
      return;
    }
    const float dr = d[0];
    if (dn > dr) return;
    int orig = -1;
    if (dn == dr) {
      orig = f2i_bits(load_pt(pts + pos).w);
      if (!(orig < f2i_bits(load_pt(pts + id[0]).w))) return;
    }
    sift_down_from_root(dn, pos, orig, k, pts, pos);
  }
  RGC_HD void sift_down_from_root(float dn, int pos, int orig, int size, const F4* pts, int lazy_pos = -1) {
    int j = 0;
    for (;;) {
      int c = 2 * j + 1;
      if (c >= size) break;
      float dc = d[c * stride];
      int ic = id[c * stride];
      if (c + 1 < size) {
        const float d2 = d[(c + 1) * stride];
        const int i2 = id[(c + 1) * stride];
        bool right_larger = d2 > dc;
        if (d2 == dc) right_larger = f2i_bits(load_pt(pts + i2).w) > f2i_bits(load_pt(pts + ic).w);
        if (right_larger) {
          c++;
          dc = d2;
          ic = i2;
        }
      }
      // stop when the larger child is not after the new entry
      bool child_after = dc > dn;
      if (dc == dn) {
        if (orig < 0 && lazy_pos >= 0) orig = f2i_bits(load_pt(pts + lazy_pos).w);
        child_after = f2i_bits(load_pt(pts + ic).w) > orig;
      }
      if (!child_after) break;
      d[j * stride] = dc;
      id[j * stride] = ic;
      j = c;
    }
    d[j * stride] = dn;
    id[j * stride] = pos;
  }
  // heap-sort in place: afterwards slots 0..cnt-1 are ascending by (d2, original index)
  RGC_HD void sort_ascending(const F4* pts) {
    for (int size = cnt; size > 1; size--) {
      const float dl = d[(size - 1) * stride];
      const int il = id[(size - 1) * stride];
      d[(size - 1) * stride] = d[0];
      id[(size - 1) * stride] = id[0];
      sift_down_from_root(dl, il, f2i_bits(load_pt(pts + il).w), size - 1, pts);
    }
  }
};

// Max-heap of packed 64-bit keys  (float bits of d2) << 32 | original index.  d2 >= +0, so the
// unsigned integer order of the key IS the lexicographic (d2, original index) order: one integer
// compare per step, no tie branches, one shared-memory word per entry.  Used by the tile kernel.
RGC_HD unsigned long long pack_key(float d2, int orig) { return ((unsigned long long)(unsigned)f2i_bits(d2) << 32) | (unsigned)orig; }
RGC_HD float key_d2(unsigned long long key) { return i2f_bits((int)(unsigned)(key >> 32)); }
struct HeapK64 {
  unsigned long long* h;
  int stride, k, cnt;
  RGC_HD void init(unsigned long long* h_, int stride_, int k_) {
    h = h_;
    stride = stride_;
    k = k_;
    cnt = 0;
  }
  RGC_HD bool full() const { return cnt == k; }
  // 4-ary max-heap: k = 20 is two levels below the root (1 + 4 + 15), so a replacement is two rounds of
  // four independent loads instead of four dependent rounds of two (the sift was 37 % of the tile
  // kernel's instructions and a chain of dependent shared-memory accesses)
  RGC_HD void sift_down(unsigned long long key, int size) {
    int j = 0;
    for (;;) {
      const int c0 = 4 * j + 1;
      if (c0 >= size) break;
      int c = c0;
      unsigned long long kc = h[c0 * stride];
#pragma unroll
      for (int u = 1; u < 4; u++)
        if (c0 + u < size) {
          const unsigned long long ku = h[(c0 + u) * stride];
          if (ku > kc) {
            kc = ku;
            c = c0 + u;
          }
        }
      if (kc <= key) break;
      h[j * stride] = kc;
      j = c;
    }
    h[j * stride] = key;
  }
  RGC_HD void insert(unsigned long long key) {
    if (cnt < k) {
      int j = cnt++;
      while (j > 0) {
        const int p = (j - 1) >> 2;
        const unsigned long long kp = h[p * stride];
        if (kp >= key) break;
        h[j * stride] = kp;
        j = p;
      }
      h[j * stride] = key;
      return;
    }
    if (key >= h[0]) return;
    sift_down(key, k);
  }
  RGC_HD void sort_ascending() {
    for (int size = cnt; size > 1; size--) {
      const unsigned long long last = h[(size - 1) * stride];
      h[(size - 1) * stride] = h[0];
      sift_down(last, size - 1);
    }
  }
};

// k = 1 container (registers): same interface as HeapK
struct Best1 {
  float d0;
  int id0;
  float lim;
  RGC_HD void reset(int, float lim_) {
    d0 = INFINITY;
    id0 = -1;
    lim = lim_;
  }
  RGC_HD bool full() const { return id0 >= 0; }
  RGC_HD float worst() const { return d0; }
  RGC_HD float bound() const { return id0 >= 0 ? fminf(lim, d0) : lim; }
  RGC_HD void insert(float dn, int pos, int orig, const F4* pts) {
    if (dn > lim || dn > d0) return;
    if (dn == d0 && id0 >= 0 && !(orig < f2i_bits(load_pt(pts + id0).w))) return;
    d0 = dn;
    id0 = pos;
  }
};

struct SearchStats {
  int nodes, lookups, candidates;
};

struct StackEntry {  // 24 bytes
  uint32_t cx_lvl;   // cx | level << 24
  uint32_t cy_mask;  // cy | child mask << 24
  uint32_t cz;
  uint32_t start, end;
  uint32_t pad;
};

template <class Top>
RGC_HD void scan_range(const GridView& g, uint32_t s, uint32_t e, float qx, float qy, float qz, Top& top, SearchStats* st) {
  for (uint32_t p = s; p < e; p++) {
    const F4 c = load_pt(g.pts + p);
    const float d2 = dist2_ref(qx, qy, qz, c.x, c.y, c.z);
    if (st) st->candidates++;
    top.insert(d2, (int)p, f2i_bits(c.w), g.pts);
  }
}

// Depth-first descent of ONE root cell (rx, ry, rz) at level lb: nearest octant first, pruning every cell whose
// (conservatively shrunk) box is farther than the current bound, scanning a cell's contiguous point range once it
// holds <= kLeafPoints points.  `stack`: kStackCap entries of per-thread scratch.
// `mkey`: the cell's Morton key (the callers build it incrementally over their root loops; rx, ry, rz are inside the grid)
template <class Top>
RGC_HD void search_root(const GridView& g, int lb, int rx, int ry, int rz, uint64_t mkey, float qx, float qy, float qz, Top& top, StackEntry* stack,
                        SearchStats* st, const CloudRange& cr) {
  if (box_dist2(g, lb, rx, ry, rz, qx, qy, qz) > top.bound()) return;
  uint32_t s, e, m;
  if (st) st->lookups++;
  if (!grid_lookup_key(g, lb, mkey | prefix_at(cr, lb), s, e, m)) return;
  int sp = 0;
  stack[sp++] = StackEntry{(uint32_t)rx | ((uint32_t)lb << 24), (uint32_t)ry | (m << 24), (uint32_t)rz, s, e, 0u};
  while (sp > 0) {
    const StackEntry n = stack[--sp];
    const int l = (int)(n.cx_lvl >> 24);
    const int cx = (int)(n.cx_lvl & 0xffffffu), cy = (int)(n.cy_mask & 0xffffffu), cz = (int)n.cz;
    const uint32_t cm = n.cy_mask >> 24;
    const float cur = top.bound();
    if (box_dist2(g, l, cx, cy, cz, qx, qy, qz) > cur) continue;
    if (st) st->nodes++;
    if (l == 0 || n.end - n.start <= (uint32_t)kLeafPoints || sp + 8 > kStackCap) {
      scan_range(g, n.start, n.end, qx, qy, qz, top, st);
      continue;
    }
    // children, nearest octant first (pushed in reverse so it is popped first)
    const float half = g.s0 * pow2f(l - 1);
    const float mx = g.ox + (float)(2 * cx + 1) * half, my = g.oy + (float)(2 * cy + 1) * half, mz = g.oz + (float)(2 * cz + 1) * half;
    const int first = (qx >= mx ? 1 : 0) | (qy >= my ? 2 : 0) | (qz >= mz ? 4 : 0);
    const uint64_t pkey = (morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz) | prefix_at(cr, l)) << 3;
    for (int j = 7; j >= 0; j--) {
      const int ci = first ^ j;
      if (!((cm >> ci) & 1u)) continue;
      const int ccx = 2 * cx + (ci & 1), ccy = 2 * cy + ((ci >> 1) & 1), ccz = 2 * cz + ((ci >> 2) & 1);
      if (box_dist2(g, l - 1, ccx, ccy, ccz, qx, qy, qz) > cur) continue;
      uint32_t cs2, ce2, cm2;
      if (st) st->lookups++;
      if (!grid_lookup_key(g, l - 1, pkey | (uint64_t)ci, cs2, ce2, cm2)) continue;  // cannot happen: mask says occupied
      stack[sp++] = StackEntry{(uint32_t)ccx | ((uint32_t)(l - 1) << 24), (uint32_t)ccy | (cm2 << 24), (uint32_t)ccz, cs2, ce2, 0u};
    }
  }
}

// root cells of the ball (q, r): the finest level whose cell edge is >= r, and the <= 3 x 3 x 3 cell range the ball
// touches at that level (hi < lo on an axis: the ball misses the grid)
struct RootRange {
  int lb, lo[3], hi[3];
};
RGC_HD RootRange root_range(const GridView& g, float qx, float qy, float qz, float bound) {
  RootRange rr;
  const int top_level = g.nlevels - 1;
  const float r = sqrtf(bound) * 1.00001f + 2.f * g.margin;  // +inf when bound is +inf
  int lb = root_level(g.s0, r, top_level);
#ifdef RGC_ROOT_SHIFT  // tuning: start the walk RGC_ROOT_SHIFT levels coarser (fewer, larger root cells)
  lb = lb + RGC_ROOT_SHIFT < top_level ? lb + RGC_ROOT_SHIFT : top_level;
#endif
  rr.lb = lb;
  const float inv_cs = g.inv_s0 * pow2f(-lb);  // == inv_s0 / 2^lb, exactly
  const int ncell = 1 << (g.nbits - lb);
  const float q3[3] = {qx, qy, qz};
  const float o3[3] = {g.ox, g.oy, g.oz};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float tl = (q3[a] - r - o3[a]) * inv_cs, th = (q3[a] + r - o3[a]) * inv_cs;
    // NaN (inf - inf) cannot occur: r = +inf gives tl = -inf, th = +inf
    int il = tl < 0.f ? 0 : (tl >= (float)ncell ? ncell : (int)tl);
    int ih = th < 0.f ? -1 : (th >= (float)ncell ? ncell - 1 : (int)th);
    // no extra slack cell: r already carries 2 margins, orders of magnitude above the rounding
    // error of this index computation
    rr.lo[a] = il;
    rr.hi[a] = ih < ncell - 1 ? ih : ncell - 1;
    if (ih < 0 || il >= ncell) rr.hi[a] = rr.lo[a] - 1;  // ball misses the grid on this axis
  }
  return rr;
}

// Exact kNN of (qx,qy,qz) in grid g into the heap `top` (call top.sort_ascending() for ordered output).
// `max_d2`: candidates with d2 > max_d2 are never needed (+inf = unbounded).  `near_pos`: a sorted
// position known to be spatially close to q (the query's own position for self-kNN), or -1.
template <class Top>
RGC_HD void knn_search(const GridView& g, float qx, float qy, float qz, int k, float max_d2, int near_pos, Top& top,
                       SearchStats* st = nullptr, const CloudRange* range = nullptr) {
  const int top_level = g.nlevels - 1;
  const CloudRange cr = range ? *range : CloudRange{0, g.n, 0ull};
  // ---- 1. initial bound from k Morton-adjacent points
  float bound = max_d2;
  if (cr.hi - cr.lo >= k) {
    int p0;
    if (near_pos >= 0) {
      p0 = near_pos - (k >> 1);
    } else {
      const int fx = cell_coord(qx, g.inv_s0, g.bias), fy = cell_coord(qy, g.inv_s0, g.bias), fz = cell_coord(qz, g.inv_s0, g.bias);
      p0 = cr.lo;
      for (int l = 0; l <= top_level; l++) {
        uint32_t s, e, m;
        if (st) st->lookups++;
        if (grid_lookup(g, l, fx >> l, fy >> l, fz >> l, s, e, m, prefix_at(cr, l))) {
          p0 = (int)s - (k >> 1) + (int)((e - s) >> 1);
          break;
        }
      }
    }
    p0 = p0 < cr.lo ? cr.lo : (p0 > cr.hi - k ? cr.hi - k : p0);
    float far = 0.f;
    for (int j = 0; j < k; j++) {
      const F4 c = load_pt(g.pts + p0 + j);
      far = fmaxf(far, dist2_ref(qx, qy, qz, c.x, c.y, c.z));
    }
    bound = fminf(bound, far);
  }
  top.reset(k, bound);

  // ---- 2. root level: finest level whose cell edge >= ball radius ; 3. depth-first descent of every root
  const RootRange rr = root_range(g, qx, qy, qz, bound);
  StackEntry stack[kStackCap];
  // (the Morton key of a root is assembled per axis: one bit-spread per loop level instead of three per cell)
  for (int rz = rr.lo[2]; rz <= rr.hi[2]; rz++) {
    const uint64_t kz = spread3((uint32_t)rz) << 2;
    for (int ry = rr.lo[1]; ry <= rr.hi[1]; ry++) {
      const uint64_t kyz = kz | (spread3((uint32_t)ry) << 1);
      for (int rx = rr.lo[0]; rx <= rr.hi[0]; rx++) search_root(g, rr.lb, rx, ry, rz, kyz | spread3((uint32_t)rx), qx, qy, qz, top, stack, st, cr);
    }
  }
}

#if defined(__CUDACC__)
// Exact 1-NN by a GROUP of G consecutive lanes (G = 4 or 8; `sub` = the lane's index in its group, `gmask` = the
// group's lanes).  The sparse-warp correspondence kernels leave G - 1 of every G lanes idle so that fewer divergent
// walks share a warp; here those lanes take a share of the SAME query instead: the ball of a hinted query touches
// up to 8 root cells, each a dependent table probe plus a short descent, and one lane walked them one after the
// other (1 760 instructions and ~10 dependent memory round trips per query, 4 of 32 lanes active: ncu of round 2).
// Lane `sub` takes the root cells sub, sub + G, ... with its own bound (the hint's distance to start with), then
// the group keeps the best (d2, original index) of its lanes.  Every root cell is still visited under a bound that
// is valid for it, so the result is the same exact nearest neighbour.  The unhinted first search also spreads
// its climb (which level holds the query's cell?) over the lanes.  All lanes of a group must call this together.
template <int G>
__device__ __forceinline__ void nn1_search_group(const GridView& g, float qx, float qy, float qz, float max_d2, int near_pos, int sub, unsigned gmask, Best1& top) {
  const int top_level = g.nlevels - 1;
  const CloudRange cr{0, g.n, 0ull};
  float bound = max_d2;
  if (g.n >= 1) {
    // every lane of the group evaluates a different seed point; the smallest distance bounds the search of all
    int p0;
    if (near_pos >= 0) {
      p0 = near_pos + sub;  // the hinted neighbour and the points that follow it in Morton order
    } else {
      const int fx = cell_coord(qx, g.inv_s0, g.bias), fy = cell_coord(qy, g.inv_s0, g.bias), fz = cell_coord(qz, g.inv_s0, g.bias);
      p0 = 0;
      for (int l0 = 0; l0 <= top_level; l0 += G) {  // lane `sub` probes level l0 + sub: the finest occupied level wins
        const int l = l0 + sub;
        uint32_t s = 0, e = 0, m;
        const bool hit = l <= top_level && grid_lookup(g, l, fx >> l, fy >> l, fz >> l, s, e, m, 0ull);
        const unsigned hits = __ballot_sync(gmask, hit) & gmask;
        if (hits) {
          const int src = __ffs(hits) - 1;  // lowest lane of the group with a hit = finest level
          const int sb = __shfl_sync(gmask, (int)s, src), eb = __shfl_sync(gmask, (int)e, src);
          p0 = sb + (int)(((long long)(eb - sb) * (2 * sub + 1)) / (2 * G));  // G points spread over that cell's range
          break;
        }
      }
    }
    p0 = p0 < 0 ? 0 : (p0 > g.n - 1 ? g.n - 1 : p0);
    const F4 c = load_pt(g.pts + p0);
    float far = dist2_ref(qx, qy, qz, c.x, c.y, c.z);
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) far = fminf(far, __shfl_xor_sync(gmask, far, o));
    bound = fminf(bound, far);
  }
  top.reset(1, bound);
  const RootRange rr = root_range(g, qx, qy, qz, bound);
  const int nx = rr.hi[0] - rr.lo[0] + 1, ny = rr.hi[1] - rr.lo[1] + 1, nz = rr.hi[2] - rr.lo[2] + 1;
  const int nroots = (nx > 0 && ny > 0 && nz > 0) ? nx * ny * nz : 0;
  StackEntry stack[kStackCap];
  for (int ri = sub; ri < nroots; ri += G) {
    const int rx = rr.lo[0] + ri % nx, ry = rr.lo[1] + (ri / nx) % ny, rz = rr.lo[2] + ri / (nx * ny);
    search_root(g, rr.lb, rx, ry, rz, morton3((uint32_t)rx, (uint32_t)ry, (uint32_t)rz), qx, qy, qz, top, stack, nullptr, cr);
  }
  // best of the group under the total order (d2, original index)
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(gmask, top.d0, o);
    const int oid = __shfl_xor_sync(gmask, top.id0, o);
    bool take = oid >= 0 && (top.id0 < 0 || od < top.d0);
    if (oid >= 0 && top.id0 >= 0 && od == top.d0 && oid != top.id0) take = f2i_bits(load_pt(g.pts + oid).w) < f2i_bits(load_pt(g.pts + top.id0).w);
    if (take) {
      top.d0 = od;
      top.id0 = oid;
    }
  }
}
#endif

}  // namespace rgc
