// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no golden vectors in the reference; Eigen's
// SelfAdjointEigenSolver / ColPivHouseholderQR and pcl::KdTreeFLANN are un-vendored, un-pinned).
//
// CPU restatement of the scan-to-map association of the mapping node,
// /root/reference/rgc_slam/src/RGC_mapping.cpp:
//   pointAssociateToMap            :1811-1820   p_map = q_w_curr * p + t_w_curr (double), stored as float
//   edge features                  :1093-1136   5-NN in the corner map; d2[4] < 1.0; centre and scatter matrix of
//                                               the 5 neighbours; line iff lambda_2 > 3 lambda_1; point_a/b = centre +- 0.1 dir
//   planar features                :1192-1240   5-NN in the surface map; d2[4] < 2.0; n = colPivHouseholderQr(A).solve(-1);
//                                               d = 1 / |n|; n /= |n|; valid iff every |n.p + d| <= 0.2
// (the "last frame" loops :1139-1189 and :1243-1290 are the same code with another pose.)
// The outputs are exactly the arguments of LidarEdgeFactor::Create / LidarPlaneNormFactor::Create.
#pragma once
#include <cmath>
#include <vector>

#include "orc_kdtree.hpp"
#include "orc_linalg.hpp"

namespace orc {

// Eigen QuaternionBase::_transformVector; q = (w, x, y, z)
inline void quat_rotate(const double q[4], const double v[3], double out[3]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  double ux = y * v[2] - z * v[1], uy = z * v[0] - x * v[2], uz = x * v[1] - y * v[0];
  ux += ux;
  uy += uy;
  uz += uz;
  out[0] = (v[0] + w * ux) + (y * uz - z * uy);
  out[1] = (v[1] + w * uy) + (z * ux - x * uz);
  out[2] = (v[2] + w * uz) + (x * uy - y * ux);
}

inline void associate_to_map(const float* p, const double q[4], const double t[3], float out[3]) {
  const double v[3] = {(double)p[0], (double)p[1], (double)p[2]};
  double r[3];
  quat_rotate(q, v, r);
  out[0] = (float)(r[0] + t[0]);
  out[1] = (float)(r[1] + t[1]);
  out[2] = (float)(r[2] + t[2]);
}

// Eigen::ColPivHouseholderQR<Matrix<double,5,3>>::solve(b) for full-rank A (rows 5, cols 3, row-major),
// pivoting on the exact remaining column norms (Eigen down-dates them; same pivots except at ties).
inline void colpiv_qr_solve_5x3(const double A_in[15], const double b_in[5], double x[3]) {
  double A[5][3], b[5];
  for (int r = 0; r < 5; r++) {
    b[r] = b_in[r];
    for (int c = 0; c < 3; c++) A[r][c] = A_in[r * 3 + c];
  }
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < 3; k++) {
    int best = k;
    double best_n = -1.0;
    for (int c = k; c < 3; c++) {
      double s = 0.0;
      for (int r = k; r < 5; r++) s += A[r][c] * A[r][c];
      if (s > best_n) {
        best_n = s;
        best = c;
      }
    }
    if (best != k) {
      for (int r = 0; r < 5; r++) std::swap(A[r][k], A[r][best]);
      std::swap(perm[k], perm[best]);
    }
    // makeHouseholderInPlace on A[k..4][k]
    double tail = 0.0;
    for (int r = k + 1; r < 5; r++) tail += A[r][k] * A[r][k];
    const double c0 = A[k][k];
    double tau = 0.0, beta = c0;
    double ess[5] = {0, 0, 0, 0, 0};
    if (tail > 2.2250738585072014e-308) {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
      for (int r = k + 1; r < 5; r++) ess[r] = A[r][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    for (int r = k + 1; r < 5; r++) A[r][k] = 0.0;
    // apply H = I - tau [1; ess][1; ess]^T to the remaining columns and to b
    for (int c = k + 1; c < 3; c++) {
      double s = A[k][c];
      for (int r = k + 1; r < 5; r++) s += ess[r] * A[r][c];
      s *= tau;
      A[k][c] -= s;
      for (int r = k + 1; r < 5; r++) A[r][c] -= s * ess[r];
    }
    double s = b[k];
    for (int r = k + 1; r < 5; r++) s += ess[r] * b[r];
    s *= tau;
    b[k] -= s;
    for (int r = k + 1; r < 5; r++) b[r] -= s * ess[r];
  }
  double y[3];
  for (int k = 2; k >= 0; k--) {
    double s = b[k];
    for (int c = k + 1; c < 3; c++) s -= A[k][c] * y[c];
    y[k] = s / A[k][k];
  }
  for (int k = 0; k < 3; k++) x[perm[k]] = y[k];
}

struct MapAssoc {
  KdTree tree;
  const float* map = nullptr;  // n x 4
  void build(const float* map_xyz1, int n) {
    map = map_xyz1;
    tree.build(map_xyz1, n);
  }
  // feats: n x 4 (x, y, z, weight); outputs: valid[n], a[n x 3], b[n x 3]
  void edges(const float* feats, int n, const double q[4], const double t[3], int* valid, double* pa, double* pb) const {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
      valid[i] = 0;
      float sel[4] = {0, 0, 0, 1.f};
      associate_to_map(feats + 4 * (size_t)i, q, t, sel);
      int idx[5];
      float d2[5];
      if (tree.knn(sel, 5, idx, d2) < 5 || !(d2[4] < 1.0)) continue;
      double c[3] = {0, 0, 0}, P[5][3];
      for (int j = 0; j < 5; j++)
        for (int a = 0; a < 3; a++) {
          P[j][a] = (double)map[4 * (size_t)idx[j] + a];
          c[a] = c[a] + P[j][a];
        }
      for (int a = 0; a < 3; a++) c[a] = c[a] / 5.0;
      double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = 0; j < 5; j++) {
        const double z[3] = {P[j][0] - c[0], P[j][1] - c[1], P[j][2] - c[2]};
        for (int r = 0; r < 3; r++)
          for (int s = 0; s < 3; s++) M[r * 3 + s] = M[r * 3 + s] + z[r] * z[s];
      }
      double ev[3], V[9];
      eigh3(M, ev, V);  // ascending, like SelfAdjointEigenSolver
      if (!(ev[2] > 3 * ev[1])) continue;
      valid[i] = 1;
      for (int a = 0; a < 3; a++) {
        const double dir = V[a * 3 + 2];
        pa[3 * (size_t)i + a] = 0.1 * dir + c[a];
        pb[3 * (size_t)i + a] = -0.1 * dir + c[a];
      }
    }
  }
  // outputs: valid[n], norm[n x 3], negative_OA_dot_norm[n]
  void planes(const float* feats, int n, const double q[4], const double t[3], int* valid, double* norm, double* dist) const {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
      valid[i] = 0;
      float sel[4] = {0, 0, 0, 1.f};
      associate_to_map(feats + 4 * (size_t)i, q, t, sel);
      int idx[5];
      float d2[5];
      if (tree.knn(sel, 5, idx, d2) < 5 || !(d2[4] < 2.0)) continue;
      double A[15], b[5] = {-1, -1, -1, -1, -1}, x[3];
      for (int j = 0; j < 5; j++)
        for (int a = 0; a < 3; a++) A[j * 3 + a] = (double)map[4 * (size_t)idx[j] + a];
      colpiv_qr_solve_5x3(A, b, x);
      const double nn = std::sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
      const double d = 1 / nn;
      const double nx = x[0] / nn, ny = x[1] / nn, nz = x[2] / nn;
      bool ok = true;
      for (int j = 0; j < 5; j++)
        if (std::fabs(((nx * A[j * 3] + ny * A[j * 3 + 1]) + nz * A[j * 3 + 2]) + d) > 0.2) {
          ok = false;
          break;
        }
      if (!ok) continue;
      valid[i] = 1;
      norm[3 * (size_t)i] = nx;
      norm[3 * (size_t)i + 1] = ny;
      norm[3 * (size_t)i + 2] = nz;
      dist[i] = d;
    }
  }
};

}  // namespace orc
