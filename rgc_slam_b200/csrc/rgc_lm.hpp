// rgc_lm.hpp — host side of the optimiser: SE(3) update, 6x6 solve, convergence test.
//
// North-star: "The Levenberg–Marquardt/Gauss–Newton 6x6 solve ... stay on the host, because
// they are tiny and sequential."  Follows LsqRegistration
// (rgc_slam/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:53-172) and so3_exp
// (rgc_slam/include/fast_gicp/so3/so3.hpp:58-77).  Matrices here are row-major 4x4 / 6x6 doubles.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace rgc {
namespace lm {

inline void mul4(const double* A, const double* B, double* C) {
  double t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) t[r * 4 + c] = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c] + A[r * 4 + 3] * B[12 + c];
  std::memcpy(C, t, sizeof(t));
}

// Solve A x = rhs for symmetric A (6x6) by LDL^T with diagonal pivoting — the factorisation
// Eigen::LDLT performs at lsq_registration_impl.hpp:111,136.  Zero pivots give zero components.
inline void solve_ldlt6(const double* A_in, const double* rhs, double* x) {
  constexpr int N = 6;
  double L[N][N], D[N], w[N];
  int p[N];
  double A[N][N];
  for (int i = 0; i < N; i++) {
    p[i] = i;
    for (int j = 0; j < N; j++) A[i][j] = A_in[i * N + j];
  }
  for (int k = 0; k < N; k++) {
    // pivot: largest remaining |diagonal| of the Schur complement
    int best = k;
    for (int i = k + 1; i < N; i++)
      if (std::fabs(A[i][i]) > std::fabs(A[best][best])) best = i;
    if (best != k) {
      std::swap(p[k], p[best]);
      for (int j = 0; j < N; j++) std::swap(A[k][j], A[best][j]);
      for (int i = 0; i < N; i++) std::swap(A[i][k], A[i][best]);
      for (int j = 0; j < k; j++) std::swap(L[k][j], L[best][j]);
    }
    D[k] = A[k][k];
    const bool ok = std::fabs(D[k]) > DBL_MIN;
    for (int i = k + 1; i < N; i++) L[i][k] = ok ? A[i][k] / D[k] : 0.0;
    for (int i = k + 1; i < N; i++)
      for (int j = k + 1; j < N; j++) A[i][j] -= L[i][k] * D[k] * L[j][k];
  }
  for (int i = 0; i < N; i++) w[i] = rhs[p[i]];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < i; j++) w[i] -= L[i][j] * w[j];
  for (int i = 0; i < N; i++) w[i] = std::fabs(D[i]) > DBL_MIN ? w[i] / D[i] : 0.0;
  for (int i = N - 1; i >= 0; i--)
    for (int j = i + 1; j < N; j++) w[i] -= L[j][i] * w[j];
  for (int i = 0; i < N; i++) x[p[i]] = w[i];
}

// delta = [so3_exp(d[0:3]).toRotationMatrix(), d[3:6]] as a 4x4 (so3.hpp:58-77 + Eigen quaternion
// to matrix).
inline void se3_delta(const double* d, double* delta) {
  const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double im, re;
  if (th2 < 1e-10) {
    const double th4 = th2 * th2;
    im = 0.5 - 1.0 / 48.0 * th2 + 1.0 / 3840.0 * th4;
    re = 1.0 - 1.0 / 8.0 * th2 + 1.0 / 384.0 * th4;
  } else {
    const double th = std::sqrt(th2);
    im = std::sin(0.5 * th) / th;
    re = std::cos(0.5 * th);
  }
  const double w = re, x = im * d[0], y = im * d[1], z = im * d[2];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
  for (int i = 0; i < 16; i++) delta[i] = 0.0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) delta[r * 4 + c] = R[r * 3 + c];
    delta[r * 4 + 3] = d[3 + r];
  }
  delta[15] = 1.0;
}

// lsq_registration_impl.hpp:82-91
inline bool is_converged(const double* delta, double rot_eps, double trans_eps) {
  double worst = 0.0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) worst = std::max(worst, std::fabs(delta[r * 4 + c] - (r == c ? 1.0 : 0.0)) * (1.0 / rot_eps));
    worst = std::max(worst, std::fabs(delta[r * 4 + 3]) * (1.0 / trans_eps));
  }
  return worst < 1;
}

}  // namespace lm
}  // namespace rgc
