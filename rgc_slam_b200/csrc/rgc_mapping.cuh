// rgc_mapping.cuh — scan-to-map association of the mapping node, included by
// rgc_gicp.cu only.  One thread per feature point (sparse lanes for small batches, like k_knn):
// pointAssociateToMap (RGC_mapping.cpp:1811-1820), exact 5-NN in the map through the same pruned
// octree walk as every other search of this library (ties by (d2, original index)), then
//   edges  (:1093-1136): centre + scatter matrix of the 5 neighbours, symmetric 3x3 eigen-solve,
//          line iff lambda_2 > 3 lambda_1, point_a / point_b = centre +- 0.1 * direction;
//   planes (:1192-1240): n = least-squares solution of A n = -1 by column-pivoted Householder QR,
//          d = 1 / |n|, n /= |n|, valid iff every |n . p_j + d| <= 0.2.
// The outputs are the arguments of LidarEdgeFactor::Create / LidarPlaneNormFactor::Create, which stay
// on the host with Ceres.
#pragma once
#include "rgc_grid.cuh"
#include "rgc_math.cuh"

namespace rgc {

struct PoseQ {
  double w, x, y, z;  // q_w_curr
  double tx, ty, tz;  // t_w_curr
};

// Eigen ColPivHouseholderQR<Matrix<double,5,3>>::solve for a full-rank system (oracle/orc_mapping.hpp)
RGC_HD void colpiv_qr_solve_5x3(double A[5][3], double b[5], double x[3]) {
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int best = k;
    double best_n = -1.0;
#pragma unroll
    for (int c = k; c < 3; c++) {
      double s = 0.0;
#pragma unroll
      for (int r = k; r < 5; r++) s += A[r][c] * A[r][c];
      if (s > best_n) {
        best_n = s;
        best = c;
      }
    }
#pragma unroll
    for (int c = k + 1; c < 3; c++)
      if (c == best) {
#pragma unroll
        for (int r = 0; r < 5; r++) {
          const double tmp = A[r][k];
          A[r][k] = A[r][c];
          A[r][c] = tmp;
        }
        const int tp = perm[k];
        perm[k] = perm[c];
        perm[c] = tp;
      }
    double tail = 0.0;
#pragma unroll
    for (int r = k + 1; r < 5; r++) tail += A[r][k] * A[r][k];
    const double c0 = A[k][k];
    double tau = 0.0, beta = c0;
    double ess[5] = {0, 0, 0, 0, 0};
    if (tail > 2.2250738585072014e-308) {
      beta = sqrt(c0 * c0 + tail);
      if (c0 >= 0.0) beta = -beta;
#pragma unroll
      for (int r = k + 1; r < 5; r++) ess[r] = A[r][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
#pragma unroll
    for (int c = k + 1; c < 3; c++) {
      double s = A[k][c];
#pragma unroll
      for (int r = k + 1; r < 5; r++) s += ess[r] * A[r][c];
      s *= tau;
      A[k][c] -= s;
#pragma unroll
      for (int r = k + 1; r < 5; r++) A[r][c] -= s * ess[r];
    }
    double s = b[k];
#pragma unroll
    for (int r = k + 1; r < 5; r++) s += ess[r] * b[r];
    s *= tau;
    b[k] -= s;
#pragma unroll
    for (int r = k + 1; r < 5; r++) b[r] -= s * ess[r];
  }
  double y[3];
  y[2] = b[2] / A[2][2];
  y[1] = (b[1] - A[1][2] * y[2]) / A[1][1];
  y[0] = (b[0] - A[0][1] * y[1] - A[0][2] * y[2]) / A[0][0];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (perm[k] == 0) x[0] = y[k];
    if (perm[k] == 1) x[1] = y[k];
    if (perm[k] == 2) x[2] = y[k];
  }
}

// pointAssociateToMap (RGC_mapping.cpp:1811-1820): q * p + t in double (Eigen _transformVector), stored as float
RGC_HD void associate_to_map(const PoseQ& T, float px, float py, float pz, float& qx, float& qy, float& qz) {
  const double vx = (double)px, vy = (double)py, vz = (double)pz;
  double ux = dsub(dmul(T.y, vz), dmul(T.z, vy)), uy = dsub(dmul(T.z, vx), dmul(T.x, vz)), uz = dsub(dmul(T.x, vy), dmul(T.y, vx));
  ux = dadd(ux, ux);
  uy = dadd(uy, uy);
  uz = dadd(uz, uz);
  qx = (float)dadd(dadd(dadd(vx, dmul(T.w, ux)), dsub(dmul(T.y, uz), dmul(T.z, uy))), T.tx);
  qy = (float)dadd(dadd(dadd(vy, dmul(T.w, uy)), dsub(dmul(T.z, ux), dmul(T.x, uz))), T.ty);
  qz = (float)dadd(dadd(dadd(vz, dmul(T.w, uz)), dsub(dmul(T.x, uy), dmul(T.y, ux))), T.tz);
}

// :1100-1133 — P: the 5 neighbours, nearest first.  Returns 1 and point_a / point_b if they form a line.
RGC_HD int edge_fit(const double P[5][3], double pa[3], double pb[3]) {
  double c[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < 5; j++) {
    c[0] += P[j][0];
    c[1] += P[j][1];
    c[2] += P[j][2];
  }
  c[0] /= 5.0;
  c[1] /= 5.0;
  c[2] /= 5.0;
  Sym3 M = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 5; j++) {
    const double zx = P[j][0] - c[0], zy = P[j][1] - c[1], zz = P[j][2] - c[2];
    M.xx += zx * zx;
    M.xy += zx * zy;
    M.xz += zx * zz;
    M.yy += zy * zy;
    M.yz += zy * zz;
    M.zz += zz * zz;
  }
  double w[3], V[3][3];
  eig_sym3(M, w, V);
  // largest and middle eigenvalue (SelfAdjointEigenSolver sorts ascending)
  int hi = 0;
  if (w[1] > w[hi]) hi = 1;
  if (w[2] > w[hi]) hi = 2;
  const int a = (hi + 1) % 3, b = (hi + 2) % 3;
  const double mid = w[a] > w[b] ? w[a] : w[b];
  if (!(w[hi] > 3 * mid)) return 0;
  for (int r = 0; r < 3; r++) {
    const double dir = hi == 0 ? V[r][0] : (hi == 1 ? V[r][1] : V[r][2]);
    pa[r] = 0.1 * dir + c[r];
    pb[r] = -0.1 * dir + c[r];
  }
  return 1;
}

// :1199-1230 — returns 1 and the unit normal / negative_OA_dot_norm if the 5 neighbours fit a plane to 0.2 m
RGC_HD int plane_fit(const double P[5][3], double nrm[3], double& dist) {
  double A[5][3], b[5] = {-1.0, -1.0, -1.0, -1.0, -1.0}, x[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < 5; j++) {
    A[j][0] = P[j][0];
    A[j][1] = P[j][1];
    A[j][2] = P[j][2];
  }
  colpiv_qr_solve_5x3(A, b, x);
  const double nn = sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
  const double d = 1 / nn;
  const double nx = x[0] / nn, ny = x[1] / nn, nz = x[2] / nn;
  int ok = 1;
  for (int j = 0; j < 5; j++)
    if (fabs(((nx * P[j][0] + ny * P[j][1]) + nz * P[j][2]) + d) > 0.2) ok = 0;
  nrm[0] = nx;
  nrm[1] = ny;
  nrm[2] = nz;
  dist = d;
  return ok;
}

#if defined(__CUDACC__)
template <bool PLANE>
__global__ void __launch_bounds__(kThreads, 4) k_map_assoc(GridView g, const unsigned char* __restrict__ feats, size_t stride, int n, int spread, PoseQ T,
                                                           int* __restrict__ valid, double* __restrict__ o1, double* __restrict__ o2) {
  extern __shared__ float heap_smem[];  // [5][kThreads] distances, then [5][kThreads] positions
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  if (gt & (spread - 1)) return;
  const int i = gt / spread;
  if (i >= n) return;
  const float* f = reinterpret_cast<const float*>(feats + (size_t)i * stride);
  float qx, qy, qz;
  associate_to_map(T, f[0], f[1], f[2], qx, qy, qz);
  HeapK top;
  top.init(heap_smem + threadIdx.x, reinterpret_cast<int*>(heap_smem + (size_t)5 * kThreads) + threadIdx.x, kThreads);
  knn_search(g, qx, qy, qz, 5, INFINITY, -1, top);
  top.sort_ascending(g.pts);
  int ok = 0;
  const float lim = PLANE ? 2.0f : 1.0f;
  if (top.cnt == 5 && top.d[4 * kThreads] < lim) {
    double P[5][3];
#pragma unroll
    for (int j = 0; j < 5; j++) {
      const F4 p = load_pt(g.pts + top.id[j * kThreads]);
      P[j][0] = (double)p.x;
      P[j][1] = (double)p.y;
      P[j][2] = (double)p.z;
    }
    if (!PLANE) {
      double pa[3], pb[3];
      ok = edge_fit(P, pa, pb);
      if (ok) {
#pragma unroll
        for (int r = 0; r < 3; r++) {
          o1[3 * (size_t)i + r] = pa[r];
          o2[3 * (size_t)i + r] = pb[r];
        }
      }
    } else {
      double nrm[3], d;
      ok = plane_fit(P, nrm, d);
      if (ok) {
        o1[3 * (size_t)i] = nrm[0];
        o1[3 * (size_t)i + 1] = nrm[1];
        o1[3 * (size_t)i + 2] = nrm[2];
        o2[i] = d;
      }
    }
  }
  valid[i] = ok;
}
#endif  // __CUDACC__

}  // namespace rgc
