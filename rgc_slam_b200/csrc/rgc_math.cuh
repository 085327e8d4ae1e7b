// rgc_math.cuh — per-point fp64 algebra of the GICP path (host/device, see rgc_common.cuh).
//
//  * covariance_from_points  : centred 3x3 covariance / k and its regularisation
//                              (fast_gicp_impl.hpp:256-293)
//  * gicp_mahalanobis        : M = (C_B + R C_A R^T)^-1          (fast_gicp_impl.hpp:146-150)
//  * gicp_point_terms        : e^T M e, H = J^T M J, b = J^T M e  (fast_gicp_impl.hpp:173-198)
//
// Symmetric 3x3 matrices are stored as 6 doubles {xx, xy, xz, yy, yz, zz}.
#pragma once
#include "rgc_common.cuh"

namespace rgc {

struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};

// 1/x and 1/sqrt(x) to ~1 ulp: hardware approximation (MUFU) + two Newton steps.  The IEEE-rounded
// double division / sqrt sequences cost ~30 instructions each and dominated the covariance kernel
// (9+ Jacobi rotations per point); results differ from IEEE by <= 1 ulp, far inside every tolerance.
RGC_HD double fast_rcp(double a) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  y = fma(y, fma(-a, y, 1.0), y);
  y = fma(y, fma(-a, y, 1.0), y);
  return y;
#else
  return 1.0 / a;
#endif
}
RGC_HD double fast_rsqrt(double a) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double h = 0.5 * a;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
#else
  return 1.0 / sqrt(a);
#endif
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (fp64).  On return A ~ V diag(w) V^T,
// V column-major-by-index: V[r][c] = component r of eigenvector c.  Unsorted.
// The reference uses Eigen::JacobiSVD on the same PSD matrix (fast_gicp_impl.hpp:273); for a
// symmetric PSD input its U and V equal the eigenvectors (up to the sign pairing handled by the
// caller), so an eigen-solver gives the same U diag(values) V^T.
RGC_HD void eig_sym3(const Sym3& A, double w[3], double V[3][3]) {
  double a[3][3] = {{A.xx, A.xy, A.xz}, {A.xy, A.yy, A.yz}, {A.xz, A.yz, A.zz}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-17 * diag || off < 1e-300) break;
#pragma unroll
    for (int pq = 0; pq < 3; pq++) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      double apq = a[p][q];
      if (fabs(apq) < 1e-300) continue;
      // rotation angle: tau = cot(2 theta), t = tan(theta) (smaller root), c = cos, s = sin
      const double tau = (a[q][q] - a[p][p]) * fast_rcp(2.0 * apq);
      const double w2 = 1.0 + tau * tau;
      const double t = (tau >= 0.0 ? 1.0 : -1.0) * fast_rcp(fabs(tau) + w2 * fast_rsqrt(w2));
      const double c = fast_rsqrt(1.0 + t * t), s = t * c;
      // A <- J^T A J with J = [[c, s], [-s, c]] on (p,q)
      a[p][p] -= t * apq;
      a[q][q] += t * apq;
      a[p][q] = a[q][p] = 0.0;
      const int r = 3 - p - q;
      double arp = a[r][p], arq = a[r][q];
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        double vip = V[i][p], viq = V[i][q];
        V[i][p] = c * vip - s * viq;
        V[i][q] = s * vip + c * viq;
      }
    }
  }
  w[0] = a[0][0];
  w[1] = a[1][1];
  w[2] = a[2][2];
}

// a * b - c * d with pinned roundings: fma(a, b, -(c * d))
RGC_HD double dprod_diff(double a, double b, double c, double d) { return dfma(a, b, -dmul(c, d)); }
// a0 * b0 + a1 * b1 + a2 * b2 with pinned roundings: fma(a2, b2, fma(a1, b1, a0 * b0))
RGC_HD double ddot3(double a0, double b0, double a1, double b1, double a2, double b2) { return dfma(a2, b2, dfma(a1, b1, dmul(a0, b0))); }

RGC_HD Sym3 inv_sym3(const Sym3& m) {
  const double c00 = dprod_diff(m.yy, m.zz, m.yz, m.yz);
  const double c01 = dprod_diff(m.xz, m.yz, m.xy, m.zz);
  const double c02 = dprod_diff(m.xy, m.yz, m.xz, m.yy);
  const double det = ddot3(m.xx, c00, m.xy, c01, m.xz, c02);
  const double id = 1.0 / det;
  Sym3 r;
  r.xx = dmul(c00, id);
  r.xy = dmul(c01, id);
  r.xz = dmul(c02, id);
  r.yy = dmul(dprod_diff(m.xx, m.zz, m.xz, m.xz), id);
  r.yz = dmul(dprod_diff(m.xy, m.xz, m.xx, m.yz), id);
  r.zz = dmul(dprod_diff(m.xx, m.yy, m.xy, m.xy), id);
  return r;
}

// sum_i val[i] * sign_i * v_i v_i^T  — the reference's U diag(values) V^T where JacobiSVD sets
// U.col(i) = sign(lambda_i) * V.col(i) (a negative eigenvalue of the PSD input can only come
// from round-off; we reproduce the resulting sign instead of "fixing" it).
RGC_HD Sym3 recompose(const double V[3][3], const double val[3], const double sgn[3]) {
  Sym3 r = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double f = val[c] * sgn[c];
    r.xx += f * V[0][c] * V[0][c];
    r.xy += f * V[0][c] * V[1][c];
    r.xz += f * V[0][c] * V[2][c];
    r.yy += f * V[1][c] * V[1][c];
    r.yz += f * V[1][c] * V[2][c];
    r.zz += f * V[2][c] * V[2][c];
  }
  return r;
}

// Unit eigenvector of the SMALLEST eigenvalue of a symmetric PSD 3x3 — all the PLANE regularisation
// needs (U diag(1,1,1e-3) V^T == I - (1 - 1e-3) n n^T).  A full Jacobi decomposition costs ~1000 fp64
// instructions per point and made k_covariance fp64-pipe bound at 17 % of HBM peak; this costs ~250:
//   lambda : Newton on the characteristic polynomial from 0 (for a PSD matrix the polynomial is
//            convex and decreasing on [0, lambda_min], so Newton climbs monotonically to the root),
//   n      : the largest cross product of two rows of (C - lambda I), polished by one inverse-iteration
//            step with the cofactor matrix of (C - lambda' I).
// Returns false (caller falls back to Jacobi) when the matrix is numerically rank <= 1 or the two
// smallest eigenvalues are too close for this shortcut to be trusted.
RGC_HD bool smallest_eigvec_sym3(const Sym3& Cin, double n[3], double& lambda_out) {
  const double tr = Cin.xx + Cin.yy + Cin.zz;
  if (!(tr > 0.0)) return false;
  const double is = fast_rcp(tr);  // scale to trace 1
  const Sym3 C = {Cin.xx * is, Cin.xy * is, Cin.xz * is, Cin.yy * is, Cin.yz * is, Cin.zz * is};
  // det(C - l I) = -l^3 + c2 l^2 - c1 l + c0
  const double c2 = 1.0;
  const double c1 = (C.xx * C.yy - C.xy * C.xy) + (C.xx * C.zz - C.xz * C.xz) + (C.yy * C.zz - C.yz * C.yz);
  const double c0 = C.xx * (C.yy * C.zz - C.yz * C.yz) - C.xy * (C.xy * C.zz - C.yz * C.xz) + C.xz * (C.xy * C.yz - C.yy * C.xz);
  double l = 0.0;
#pragma unroll 1
  for (int it = 0; it < 12; it++) {
    const double f = ((-l + c2) * l - c1) * l + c0;
    const double fp = (-3.0 * l + 2.0 * c2) * l - c1;
    if (!(fp < 0.0)) break;
    const double step = f * fast_rcp(fp);
    l -= step;
    if (fabs(step) <= 1e-17) break;
  }
  if (!(l > -1e-12) || !(l < 0.34)) return false;
  // the second eigenvalue must be clearly separated: l2 + l1 = 1 - l, l1 l2 = c0 / l (or via c1)
  // => gap test on the reduced quadratic  m^2 - (1 - l) m + (c1 - l (1 - l)) = 0
  const double sum = 1.0 - l, prod = c1 - l * sum;
  const double disc = sum * sum - 4.0 * prod;
  const double l2 = 0.5 * (sum - sqrt(disc > 0.0 ? disc : 0.0));
  if (!(l2 - l > 1e-7)) return false;  // near-degenerate pair: let Jacobi decide
  auto eigvec = [&](double mu, double v[3]) {
    const double r0[3] = {C.xx - mu, C.xy, C.xz}, r1[3] = {C.xy, C.yy - mu, C.yz}, r2[3] = {C.xz, C.yz, C.zz - mu};
    const double a[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    const double b[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    const double c[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double na = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], nb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2], nc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double* best = a;
    double nbest = na;
    if (nb > nbest) { best = b; nbest = nb; }
    if (nc > nbest) { best = c; nbest = nc; }
    const double inv = fast_rsqrt(nbest);
    v[0] = best[0] * inv; v[1] = best[1] * inv; v[2] = best[2] * inv;
    return nbest;
  };
  double v[3];
  const double nb0 = eigvec(l, v);
  if (!(nb0 > 1e-24)) return false;  // (C - l I) numerically rank <= 1
  // one inverse-iteration polish: y = adj(C - mu I) v with mu just below l (adjugate ~ n n^T / gap)
  {
    const double mu = l - 1e-9 * (l2 - l) - 1e-18;
    const double a = C.xx - mu, b = C.yy - mu, c = C.zz - mu;
    const double A00 = b * c - C.yz * C.yz, A01 = C.xz * C.yz - C.xy * c, A02 = C.xy * C.yz - C.xz * b;
    const double A11 = a * c - C.xz * C.xz, A12 = C.xy * C.xz - a * C.yz, A22 = a * b - C.xy * C.xy;
    const double y0 = A00 * v[0] + A01 * v[1] + A02 * v[2], y1 = A01 * v[0] + A11 * v[1] + A12 * v[2], y2 = A02 * v[0] + A12 * v[1] + A22 * v[2];
    const double ny = y0 * y0 + y1 * y1 + y2 * y2;
    if (ny > 1e-60) {
      const double inv = fast_rsqrt(ny);
      v[0] = y0 * inv; v[1] = y1 * inv; v[2] = y2 * inv;
    }
  }
  n[0] = v[0]; n[1] = v[1]; n[2] = v[2];
  lambda_out = l * tr;
  return true;
}

// fast_gicp_impl.hpp:264-293, every method through the full eigen-decomposition.  Kept out of line:
// on the device it is the rare path (non-PLANE methods, degenerate neighbourhoods) and inlining it
// into k_covariance cost that kernel a third of its occupancy in registers.
RGC_HD_NOINLINE Sym3 regularize_cov_general(const Sym3& cov, int method) {
  if (method == REG_FROBENIUS) {
    const double lambda = 1e-3;
    Sym3 C = cov;
    C.xx += lambda;
    C.yy += lambda;
    C.zz += lambda;
    Sym3 Ci = inv_sym3(C);
    double nrm = sqrt(Ci.xx * Ci.xx + Ci.yy * Ci.yy + Ci.zz * Ci.zz + 2.0 * (Ci.xy * Ci.xy + Ci.xz * Ci.xz + Ci.yz * Ci.yz));
    double s = 1.0 / nrm;
    Sym3 Cn = {Ci.xx * s, Ci.xy * s, Ci.xz * s, Ci.yy * s, Ci.yz * s, Ci.zz * s};
    return inv_sym3(Cn);
  }
  double w[3], V[3][3];
  eig_sym3(cov, w, V);
  // singular values = |eigenvalues|; rank them descending like JacobiSVD's final sort
  double sv[3] = {fabs(w[0]), fabs(w[1]), fabs(w[2])};
  double sgn[3] = {w[0] < 0 ? -1.0 : 1.0, w[1] < 0 ? -1.0 : 1.0, w[2] < 0 ? -1.0 : 1.0};
  int rank[3];  // rank[c] = position of column c in descending order
#pragma unroll
  for (int c = 0; c < 3; c++) {
    int r = 0;
#pragma unroll
    for (int o = 0; o < 3; o++)
      if (o != c && (sv[o] > sv[c] || (sv[o] == sv[c] && o < c))) r++;
    rank[c] = r;
  }
  double val[3];
  if (method == REG_PLANE) {
#pragma unroll
    for (int c = 0; c < 3; c++) val[c] = (rank[c] == 2) ? 1e-3 : 1.0;
  } else if (method == REG_MIN_EIG) {
#pragma unroll
    for (int c = 0; c < 3; c++) val[c] = sv[c] > 1e-3 ? sv[c] : 1e-3;
  } else {  // REG_NORMALIZED_MIN_EIG
    double mx = sv[0] > sv[1] ? (sv[0] > sv[2] ? sv[0] : sv[2]) : (sv[1] > sv[2] ? sv[1] : sv[2]);
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double v = sv[c] / mx;
      val[c] = v > 1e-3 ? v : 1e-3;
    }
  }
  return recompose(V, val, sgn);
}

RGC_HD Sym3 regularize_cov(const Sym3& cov, int method) {
  if (method == REG_NONE) return cov;
  if (method == REG_PLANE) {
    double n[3], lam;
    if (smallest_eigvec_sym3(cov, n, lam)) {
      // I - (1 - s 1e-3) n n^T, s = sign pairing of the reference's SVD (negative only through round-off)
      const double f = 1.0 - (lam < 0.0 ? -1e-3 : 1e-3);
      return Sym3{1.0 - f * n[0] * n[0], -f * n[0] * n[1], -f * n[0] * n[2], 1.0 - f * n[1] * n[1], -f * n[1] * n[2], 1.0 - f * n[2] * n[2]};
    }
  }
  return regularize_cov_general(cov, method);
}

// Row-major 3x4 rigid transform in double.
struct Rt {
  double m[12];
};

RGC_HD void transform_d(const Rt& T, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = dadd(ddot3(T.m[0], x, T.m[1], y, T.m[2], z), T.m[3]);
  oy = dadd(ddot3(T.m[4], x, T.m[5], y, T.m[6], z), T.m[7]);
  oz = dadd(ddot3(T.m[8], x, T.m[9], y, T.m[10], z), T.m[11]);
}

// M = (C_B + R C_A R^T)^-1 (the 4x4 of the reference is block diagonal with a pinned 1, so its
// inverse is this 3x3 inverse; fast_gicp_impl.hpp:146-150)
RGC_HD Sym3 gicp_mahalanobis(const Rt& T, const Sym3& CA, const Sym3& CB) {
  const double* R = T.m;
  double RC[3][3];
  const double A[3][3] = {{CA.xx, CA.xy, CA.xz}, {CA.xy, CA.yy, CA.yz}, {CA.xz, CA.yz, CA.zz}};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) RC[i][j] = ddot3(R[i * 4 + 0], A[0][j], R[i * 4 + 1], A[1][j], R[i * 4 + 2], A[2][j]);
  Sym3 S;
  S.xx = dadd(CB.xx, ddot3(RC[0][0], R[0], RC[0][1], R[1], RC[0][2], R[2]));
  S.xy = dadd(CB.xy, ddot3(RC[0][0], R[4], RC[0][1], R[5], RC[0][2], R[6]));
  S.xz = dadd(CB.xz, ddot3(RC[0][0], R[8], RC[0][1], R[9], RC[0][2], R[10]));
  S.yy = dadd(CB.yy, ddot3(RC[1][0], R[4], RC[1][1], R[5], RC[1][2], R[6]));
  S.yz = dadd(CB.yz, ddot3(RC[1][0], R[8], RC[1][1], R[9], RC[1][2], R[10]));
  S.zz = dadd(CB.zz, ddot3(RC[2][0], R[8], RC[2][1], R[9], RC[2][2], R[10]));
  return inv_sym3(S);
}

// Accumulator layout of one linearize() reduction: 28 doubles + 1 inlier count
//   [0]      sum e^T M e
//   [1..21]  upper triangle of H (6x6), row-major: (0,0)(0,1)..(0,5)(1,1)..(5,5)
//   [22..27] b
constexpr int kAccN = 28;

RGC_HD double gicp_error_term(const Rt& T, const Sym3& M, float px, float py, float pz, float qx, float qy, float qz) {
  double ax, ay, az;
  transform_d(T, f2d(px), f2d(py), f2d(pz), ax, ay, az);
  const double ex = dsub(f2d(qx), ax), ey = dsub(f2d(qy), ay), ez = dsub(f2d(qz), az);
  const double mx = ddot3(M.xx, ex, M.xy, ey, M.xz, ez);
  const double my = ddot3(M.xy, ex, M.yy, ey, M.yz, ez);
  const double mz = ddot3(M.xz, ex, M.yz, ey, M.zz, ez);
  return ddot3(ex, mx, ey, my, ez, mz);
}

// sum_i J[i][c] * v_i for J = [skew(a) | -I] (3x6): the only products that are not 0 or -v survive, with
// pinned roundings.  c is a compile-time constant at every call site (fully unrolled loops).
RGC_HD double jt_dot(int c, const double a[3], double v0, double v1, double v2) {
  switch (c) {
    case 0: return dprod_diff(a[2], v1, a[1], v2);   //  a2 v1 - a1 v2
    case 1: return dprod_diff(a[0], v2, a[2], v0);   //  a0 v2 - a2 v0
    case 2: return dprod_diff(a[1], v0, a[0], v1);   //  a1 v0 - a0 v1
    case 3: return -v0;
    case 4: return -v1;
    default: return -v2;
  }
}

// acc += terms of one correspondence.  J = [skew(a) | -I], a = T p.
RGC_HD void gicp_point_terms(const Rt& T, const Sym3& M, float px, float py, float pz, float qx, float qy, float qz, double* acc) {
  double a[3];
  transform_d(T, f2d(px), f2d(py), f2d(pz), a[0], a[1], a[2]);
  const double e[3] = {dsub(f2d(qx), a[0]), dsub(f2d(qy), a[1]), dsub(f2d(qz), a[2])};
  const double Mm[3][3] = {{M.xx, M.xy, M.xz}, {M.xy, M.yy, M.yz}, {M.xz, M.yz, M.zz}};
  double Me[3];
#pragma unroll
  for (int i = 0; i < 3; i++) Me[i] = ddot3(Mm[i][0], e[0], Mm[i][1], e[1], Mm[i][2], e[2]);
  acc[0] = dadd(acc[0], ddot3(e[0], Me[0], e[1], Me[1], e[2], Me[2]));
  // MJ = M J (3x6); H = J^T (M J), upper triangle; b = J^T (M e)
  double MJ[3][6];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int c = 0; c < 6; c++) MJ[i][c] = jt_dot(c, a, Mm[i][0], Mm[i][1], Mm[i][2]);
  int o = 1;
#pragma unroll
  for (int r = 0; r < 6; r++)
#pragma unroll
    for (int c = r; c < 6; c++) {
      acc[o] = dadd(acc[o], jt_dot(r, a, MJ[0][c], MJ[1][c], MJ[2][c]));
      o++;
    }
#pragma unroll
  for (int r = 0; r < 6; r++) acc[22 + r] = dadd(acc[22 + r], jt_dot(r, a, Me[0], Me[1], Me[2]));
}

// Centred covariance of `found` gathered points (k columns; missing columns are zero, matching
// the oracle's handling of k > N), divided by k.  fast_gicp_impl.hpp:256-262.
template <class GetPt>
RGC_HD Sym3 covariance_from_points(int found, int k, GetPt get) {
  // one pass: moments of (p_j - p_0) in fp64 (p_0 = the query itself, so the shifted data is as
  // small as the neighbourhood and E[dd^T] - E[d]E[d]^T loses nothing), then re-centre.
  if (found <= 0) return Sym3{0, 0, 0, 0, 0, 0};
  const F4 p0 = get(0);
  const double ox = f2d(p0.x), oy = f2d(p0.y), oz = f2d(p0.z);
  double sx = 0.0, sy = 0.0, sz = 0.0;
  Sym3 c = {0, 0, 0, 0, 0, 0};
  for (int j = 1; j < found; j++) {
    const F4 p = get(j);
    const double dx = f2d(p.x) - ox, dy = f2d(p.y) - oy, dz = f2d(p.z) - oz;
    sx += dx;
    sy += dy;
    sz += dz;
    c.xx += dx * dx;
    c.xy += dx * dy;
    c.xz += dx * dz;
    c.yy += dy * dy;
    c.yz += dy * dz;
    c.zz += dz * dz;
  }
  // missing columns (k > N) are zero vectors in the reference's matrix, i.e. the point -p_0 in
  // shifted coordinates; they are rare (tiny clouds) and handled exactly
  const int missing = k - found;
  if (missing > 0) {
    const double m = (double)missing;
    sx -= m * ox;
    sy -= m * oy;
    sz -= m * oz;
    c.xx += m * ox * ox;
    c.xy += m * ox * oy;
    c.xz += m * ox * oz;
    c.yy += m * oy * oy;
    c.yz += m * oy * oz;
    c.zz += m * oz * oz;
  }
  const double ik = 1.0 / (double)k;
  const double mx = sx * ik, my = sy * ik, mz = sz * ik;
  c.xx = c.xx * ik - mx * mx;
  c.xy = c.xy * ik - mx * my;
  c.xz = c.xz * ik - mx * mz;
  c.yy = c.yy * ik - my * my;
  c.yz = c.yz * ik - my * mz;
  c.zz = c.zz * ik - mz * mz;
  return c;
}

}  // namespace rgc
