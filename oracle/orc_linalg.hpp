// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use anything under oracle/.  PARITY UNPINNED: the reference ships no golden vectors
//  and cannot be compiled here (no PCL/Eigen/FLANN).
//
// Small dense linear algebra the reference obtains from Eigen (un-vendored, version not
// pinned by the reference; README.md:24-31 says "Eigen >= 3.3.4").  Each routine restates the
// *published* Eigen 3.3 algorithm named in its comment so that degenerate-case behaviour
// (U != V columns, sign flips) follows the reference's dependency as closely as we can know.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace orc {

// ---- 3x3 helpers (row-major double[9]) -------------------------------------------------
inline void mat3_mul(const double* A, const double* B, double* C) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += A[i * 3 + k] * B[k * 3 + j];
      T[i * 3 + j] = s;
    }
  std::memcpy(C, T, sizeof(T));
}

// Jacobi plane rotation (c,s) as in Eigen::JacobiRotation.
struct JRot {
  double c, s;
};

// Eigen::JacobiRotation::makeJacobi(x, y, z) for the real symmetric 2x2 [[x,y],[y,z]]
// (Eigen/src/Jacobi/Jacobi.h).
inline bool make_jacobi(double x, double y, double z, JRot* r) {
  double deno = 2.0 * std::fabs(y);
  if (deno < DBL_MIN) {
    r->c = 1.0;
    r->s = 0.0;
    return false;
  }
  double tau = (x - z) / deno;
  double w = std::sqrt(tau * tau + 1.0);
  double t = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  double sign_t = t > 0 ? 1.0 : -1.0;
  double n = 1.0 / std::sqrt(t * t + 1.0);
  r->s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r->c = n;
  return true;
}

// M.applyOnTheLeft(p,q,j): rows p,q of M <- [c s; -s c]^T-style as Eigen defines:
//   x_i' =  c x_i + s y_i ;  y_i' = -s x_i + c y_i   with x=row p, y=row q  (j.adjoint applied)
// Eigen: applyOnTheLeft(p,q,j) does apply_rotation_in_the_plane(row p, row q, j)
//   which computes  x' = c x + s y ; y' = -s x + c y   (for real scalars).
inline void rot_left(double* M, int n, int p, int q, const JRot& j) {
  for (int i = 0; i < n; i++) {
    double x = M[p * n + i], y = M[q * n + i];
    M[p * n + i] = j.c * x + j.s * y;
    M[q * n + i] = -j.s * x + j.c * y;
  }
}
// M.applyOnTheRight(p,q,j): apply_rotation_in_the_plane(col p, col q, j.transpose())
//   x' = c x - s y ; y' = s x + c y
inline void rot_right(double* M, int n, int p, int q, const JRot& j) {
  for (int i = 0; i < n; i++) {
    double x = M[i * n + p], y = M[i * n + q];
    M[i * n + p] = j.c * x - j.s * y;
    M[i * n + q] = j.s * x + j.c * y;
  }
}

// Eigen::internal::real_2x2_jacobi_svd (Eigen/src/misc/RealSvd2x2.h).
inline void real_2x2_jacobi_svd(const double* W, int n, int p, int q, JRot* j_left, JRot* j_right) {
  double m[4] = {W[p * n + p], W[p * n + q], W[q * n + p], W[q * n + q]};
  JRot rot1;
  double t = m[0] + m[3];
  double d = m[2] - m[1];
  if (std::fabs(d) < DBL_MIN) {
    rot1.s = 0.0;
    rot1.c = 1.0;
  } else {
    double u = t / d;
    double tmp = std::sqrt(1.0 + u * u);
    rot1.s = 1.0 / tmp;
    rot1.c = u / tmp;
  }
  rot_left(m, 2, 0, 1, rot1);
  make_jacobi(m[0], m[1], m[3], j_right);
  // *j_left = rot1 * j_right->transpose();   (JacobiRotation product)
  JRot jt = {j_right->c, -j_right->s};
  j_left->c = rot1.c * jt.c - rot1.s * jt.s;
  j_left->s = rot1.c * jt.s + rot1.s * jt.c;
}

// Eigen::JacobiSVD<Matrix3d>(A, ComputeFullU|ComputeFullV) — two-sided Jacobi, square case
// (Eigen/src/SVD/JacobiSVD.h, compute()).  A, U, V row-major 3x3; sv descending.
inline void jacobi_svd3(const double* A, double* U, double* sv, double* V) {
  const int n = 3;
  const double precision = 2.0 * DBL_EPSILON;
  const double considerAsZero = DBL_MIN;
  double scale = 0.0;
  for (int i = 0; i < 9; i++) scale = std::max(scale, std::fabs(A[i]));
  if (scale == 0.0) scale = 1.0;
  double W[9];
  for (int i = 0; i < 9; i++) W[i] = A[i] / scale;
  for (int i = 0; i < 9; i++) U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  double maxDiag = 0.0;
  for (int i = 0; i < n; i++) maxDiag = std::max(maxDiag, std::fabs(W[i * n + i]));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 1000) {
    finished = true;
    for (int p = 1; p < n; ++p) {
      for (int q = 0; q < p; ++q) {
        double threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::fabs(W[p * n + q]) > threshold || std::fabs(W[q * n + p]) > threshold) {
          finished = false;
          JRot jl, jr;
          real_2x2_jacobi_svd(W, n, p, q, &jl, &jr);
          rot_left(W, n, p, q, jl);
          JRot jlt = {jl.c, -jl.s};
          rot_right(U, n, p, q, jlt);
          rot_right(W, n, p, q, jr);
          rot_right(V, n, p, q, jr);
          maxDiag = std::max(maxDiag, std::max(std::fabs(W[p * n + p]), std::fabs(W[q * n + q])));
        }
      }
    }
  }
  for (int i = 0; i < n; i++) {
    double a = std::fabs(W[i * n + i]);
    sv[i] = a;
    if (a != 0.0) {
      double f = W[i * n + i] / a;
      for (int r = 0; r < n; r++) U[r * n + i] *= f;
    }
  }
  for (int i = 0; i < n; i++) sv[i] *= scale;
  // sort descending, swapping columns
  for (int i = 0; i < n; i++) {
    int pos = i;
    double best = sv[i];
    for (int j = i + 1; j < n; j++)
      if (sv[j] > best) {
        best = sv[j];
        pos = j;
      }
    if (best == 0.0) break;
    if (pos != i) {
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < n; r++) {
        std::swap(U[r * n + i], U[r * n + pos]);
        std::swap(V[r * n + i], V[r * n + pos]);
      }
    }
  }
}

// Symmetric 3x3 eigen-decomposition, eigenvalues ascending, columns of V normalised.
// Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d> (scanRegistration.cpp:371), whose
// tridiagonal-QL iteration is mathematically equivalent; eigenvector SIGNS are implementation
// defined in Eigen, so callers compare vectors up to sign.  Cyclic Jacobi here.
inline void eigh3(const double* A, double* evals, double* V) {
  double W[9];
  std::memcpy(W, A, sizeof(W));
  for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = std::fabs(W[1]) + std::fabs(W[2]) + std::fabs(W[5]);
    double diag = std::fabs(W[0]) + std::fabs(W[4]) + std::fabs(W[8]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        JRot j;
        if (!make_jacobi(W[p * 3 + p], W[p * 3 + q], W[q * 3 + q], &j)) continue;
        JRot jt = {j.c, -j.s};
        rot_left(W, 3, p, q, jt);
        rot_right(W, 3, p, q, j);
        rot_right(V, 3, p, q, j);
        W[q * 3 + p] = W[p * 3 + q] = 0.5 * (W[q * 3 + p] + W[p * 3 + q]);
      }
  }
  int order[3] = {0, 1, 2};
  double ev[3] = {W[0], W[4], W[8]};
  std::sort(order, order + 3, [&](int a, int b) { return ev[a] < ev[b]; });
  double Vs[9];
  for (int c = 0; c < 3; c++) {
    evals[c] = ev[order[c]];
    for (int r = 0; r < 3; r++) Vs[r * 3 + c] = V[r * 3 + order[c]];
  }
  std::memcpy(V, Vs, sizeof(Vs));
}

// General 4x4 inverse by cofactors — what Eigen::Matrix4d::inverse() does
// (Eigen/src/LU/InverseImpl.h compute_inverse_size4: cofactor expansion / determinant).
inline void inverse4(const double* m, double* inv) {
  double t[16];
  t[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  t[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  t[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  t[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  t[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  t[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  t[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  t[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  t[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  t[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  t[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  t[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  t[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  t[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  t[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  t[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * t[0] + m[1] * t[4] + m[2] * t[8] + m[3] * t[12];
  double id = 1.0 / det;
  for (int i = 0; i < 16; i++) inv[i] = t[i] * id;
}

inline void inverse3(const double* m, double* inv) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c01 = m[5] * m[6] - m[3] * m[8];
  double c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  inv[0] = c00 * id;
  inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  inv[3] = c01 * id;
  inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  inv[6] = c02 * id;
  inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// Eigen::LDLT<Matrix<double,6,6>>(A).solve(rhs): symmetric-pivoted LDL^T (lower), then
// P, L^-1, D^-1 (entries with |d| <= DBL_MIN give 0), L^-T, P^T.  (Eigen/src/Cholesky/LDLT.h)
inline void ldlt6_solve(const double* Ain, const double* rhs, double* x) {
  const int n = 6;
  double A[36];
  std::memcpy(A, Ain, sizeof(A));
  int perm[6];
  for (int i = 0; i < n; i++) perm[i] = i;
  int transp[6];
  for (int k = 0; k < n; k++) {
    int piv = k;
    double big = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i * n + i]) > big) {
        big = std::fabs(A[i * n + i]);
        piv = i;
      }
    transp[k] = piv;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[piv * n + j]);
      for (int i = 0; i < n; i++) std::swap(A[i * n + k], A[i * n + piv]);
    }
    // lower part holds L (unit) and D on the diagonal
    for (int j = 0; j < k; j++) A[k * n + k] -= A[k * n + j] * A[k * n + j] * A[j * n + j];
    double dk = A[k * n + k];
    for (int i = k + 1; i < n; i++) {
      double s = A[i * n + k];
      for (int j = 0; j < k; j++) s -= A[i * n + j] * A[k * n + j] * A[j * n + j];
      A[i * n + k] = (std::fabs(dk) > DBL_MIN) ? s / dk : 0.0;
    }
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = rhs[i];
  for (int k = 0; k < n; k++) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) y[i] -= A[i * n + j] * y[j];
  for (int i = 0; i < n; i++) {
    double d = A[i * n + i];
    y[i] = (std::fabs(d) > DBL_MIN) ? y[i] / d : 0.0;
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = i + 1; j < n; j++) y[i] -= A[j * n + i] * y[j];
  for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[transp[k]]);
  for (int i = 0; i < n; i++) x[i] = y[i];
  (void)perm;
}

// fast_gicp so3_exp (rgc_slam/include/fast_gicp/so3/so3.hpp:58-77) -> quaternion (w,x,y,z),
// then Eigen::Quaterniond::toRotationMatrix() (row-major R).
inline void so3_exp_quat(const double* omega, double* q) {
  double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    double theta = std::sqrt(theta_sq);
    double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  q[0] = real_factor;
  q[1] = imag_factor * omega[0];
  q[2] = imag_factor * omega[1];
  q[3] = imag_factor * omega[2];
}

inline void quat_to_rot(const double* q, double* R) {
  // Eigen::QuaternionBase::toRotationMatrix (no normalisation)
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}

}  // namespace orc
