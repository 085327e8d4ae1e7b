// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (the reference holds no golden vectors for
// this path and pcl::VoxelGrid lives in PCL, which is neither vendored nor version-pinned by the
// reference: README.md:24-31 names ROS Melodic/Noetic, i.e. PCL 1.8 / 1.10).
//
// CPU restatement of the step in front of the registration:
//   * adjustDistortion, /root/reference/rgc_slam/src/RGC_odometer.cpp:1441-1481 (third loop, the full
//     cloud): s = 1 - frac(intensity) / SCAN_PERIOD in FLOAT (SCAN_PERIOD is `const float`, :323),
//     q = Identity.slerp(s, q_last_curr^-1) (Eigen 3.3 QuaternionBase::slerp), t = s * t_last_curr,
//     p' = q * (p - t) (Eigen QuaternionBase::_transformVector), stored back into float fields;
//   * pcl::VoxelGrid<PointXYZI>::applyFilter with the defaults the call sites use
//     (RGC_odometer.cpp:975-991: leaf 0.2 / 0.3 m, downsample_all_data = true, min_points_per_voxel = 0),
//     restated from PCL's published algorithm (filters/impl/voxel_grid.hpp): bounding box -> min_b /
//     div_b in float arithmetic, voxel index = ijk . (1, dx, dx*dy), sort by index, per-voxel
//     CentroidPoint (float sums of x, y, z, intensity, divided by float(n)), output in index order.
//     PCL sorts with std::sort (order of equal keys unspecified); this restatement and the GPU path
//     both add a voxel's points in ascending input order, so a PCL run may differ in the last ulp.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace orc {

// xyzi: n x 4 floats (x, y, z, intensity); q = (w, x, y, z) of q_last_curr; out: n x 4
inline void deskew(const float* xyzi, int n, const double q_last_curr[4], const double t_last_curr[3], float scan_period, float* out) {
  // Eigen Quaternion::inverse(): conjugate / squaredNorm (coeff order x, y, z, w in the dot product)
  const double qw = q_last_curr[0], qx = q_last_curr[1], qy = q_last_curr[2], qz = q_last_curr[3];
  const double n2 = ((qx * qx + qy * qy) + qz * qz) + qw * qw;
  double iw = 0, ix = 0, iy = 0, iz = 0;
  if (n2 > 0.0) {
    iw = qw / n2;
    ix = -qx / n2;
    iy = -qy / n2;
    iz = -qz / n2;
  }
  const double one = 1.0 - std::numeric_limits<double>::epsilon();
  for (int i = 0; i < n; i++) {
    const float* p = xyzi + 4 * (size_t)i;
    const float inten = p[3];
    const float sf = 1 - (inten - int(inten)) / scan_period;  // float arithmetic, as in the reference
    const double s = sf;
    // Identity.slerp(s, q_inv)
    const double d = iw;  // 0*x + 0*y + 0*z + 1*w
    const double absd = std::fabs(d);
    double scale0, scale1;
    if (absd >= one) {
      scale0 = 1.0 - s;
      scale1 = s;
    } else {
      const double theta = std::acos(absd);
      const double sin_theta = std::sin(theta);
      scale0 = std::sin((1.0 - s) * theta) / sin_theta;
      scale1 = std::sin(s * theta) / sin_theta;
    }
    if (d < 0.0) scale1 = -scale1;
    const double w = scale0 * 1.0 + scale1 * iw, x = scale0 * 0.0 + scale1 * ix, y = scale0 * 0.0 + scale1 * iy, z = scale0 * 0.0 + scale1 * iz;
    const double vx = (double)p[0] - s * t_last_curr[0], vy = (double)p[1] - s * t_last_curr[1], vz = (double)p[2] - s * t_last_curr[2];
    // _transformVector: uv = vec x v; uv += uv; v + w * uv + vec x uv
    double ux = y * vz - z * vy, uy = z * vx - x * vz, uz = x * vy - y * vx;
    ux += ux;
    uy += uy;
    uz += uz;
    const double rx = (vx + w * ux) + (y * uz - z * uy), ry = (vy + w * uy) + (z * ux - x * uz), rz = (vz + w * uz) + (x * uy - y * ux);
    float* o = out + 4 * (size_t)i;
    o[0] = (float)rx;
    o[1] = (float)ry;
    o[2] = (float)rz;
    o[3] = inten;
  }
}

// returns the number of output points; out (capacity >= n x 4) in ascending voxel-index order.
// `passthrough` is set when the index space overflows int32 and PCL returns the input unchanged.
inline int voxel_grid(const float* xyzi, int n, float leaf, float* out, int* passthrough) {
  *passthrough = 0;
  if (n == 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      const float v = xyzi[4 * (size_t)i + a];
      mn[a] = std::min(mn[a], v);
      mx[a] = std::max(mx[a], v);
    }
  int64_t d[3];
  for (int a = 0; a < 3; a++) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
  if (d[0] * d[1] * d[2] > (int64_t)std::numeric_limits<int32_t>::max()) {
    std::copy(xyzi, xyzi + 4 * (size_t)n, out);
    *passthrough = 1;
    return n;
  }
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; a++) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    const int max_b = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<unsigned, int>> iv((size_t)n);
  for (int i = 0; i < n; i++) {
    const float* p = xyzi + 4 * (size_t)i;
    const int i0 = (int)(std::floor(p[0] * inv) - (float)min_b[0]);
    const int i1 = (int)(std::floor(p[1] * inv) - (float)min_b[1]);
    const int i2 = (int)(std::floor(p[2] * inv) - (float)min_b[2]);
    iv[i] = {(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), i};
  }
  std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
  int m = 0;
  for (size_t s = 0; s < iv.size();) {
    size_t e = s + 1;
    while (e < iv.size() && iv[e].first == iv[s].first) e++;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (size_t j = s; j < e; j++) {
      const float* p = xyzi + 4 * (size_t)iv[j].second;
      sx += p[0];
      sy += p[1];
      sz += p[2];
      si += p[3];
    }
    const float cnt = (float)(e - s);
    float* o = out + 4 * (size_t)m;
    o[0] = sx / cnt;
    o[1] = sy / cnt;
    o[2] = sz / cnt;
    o[3] = si / cnt;
    m++;
    s = e;
  }
  return m;
}

}  // namespace orc
