// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.hpp header).  PARITY UNPINNED.
//
// CPU restatement of the numeric body of ScanRegistration::laserCloudHandler
//   rgc_slam/src/scanRegistration.cpp:110-663  and removeClosedPointCloud :732-763.
// ROS I/O (:104-109, :687-727) is out of scope.  The reference keeps its per-point arrays as
// never-cleared members (scanRegistration.cpp:42-52); entries it does not rewrite each frame
// are stale there.  This restatement zero-initialises every array per call — all reads the
// reference performs on stale entries are masked (`&& range_vec[i] < 2`, :259,:282) or fall
// outside every segment, so results on [0, cloudSize) are unaffected.
// Float/double mixing follows the reference expression by expression.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_linalg.hpp"

namespace orc {

struct FeatureParams {
  int n_scans = 16;               // scanRegistration.cpp:29
  double minimum_range = 0.5;     // :31, launch default
  double maximum_range = 80.0;    // run.launch maxmum_range
  int use_intensity = 1;          // :30
};

struct FeatureOut {
  int cloud_size = 0;                   // points kept after ring assignment
  std::vector<float> cloud;             // ring-ordered x,y,z,intensity(=scanID+0.1*relTime), 4 floats
  std::vector<int> src_index;           // index into the *input* cloud of each ordered point
  std::vector<int> intensity_num;       // smoothed integer intensity (:257-268)
  std::vector<int> scan_start, scan_end;  // :223,:229
  std::vector<float> range_vec, scan_angle, curvature, inten_curvature, curvature2, distance_source, other_source;
  std::vector<int> neighbor_picked, inten_neighbor_picked, label, inten_label, ground_marked;
  // compacted outputs, in the reference's push_back order (indices into the ordered cloud)
  std::vector<int> corner_sharp, corner_less_sharp, surf_flat, surf_less_flat, inten_sharp, inten_less_sharp, ground_points;
  std::vector<float> corner_sharp_w, surf_flat_w, inten_sharp_w;  // normal_x weights (:501,:554,:609)
  int inten_merged = 0;  // 1 if intenPointsSharp was appended to cornerPointsSharp (:653-656)
  int ground_size = 0;
  double groundparam[11] = {0};  // normx,y,z, vector1 x,y,z, vector2 x,y,z, distance, source (:420-430)
  double ground_evals[3] = {0};
};

static inline float absf(float v) { return v < 0 ? -v : v; }

inline void extract_features(const float* in_xyzi, int n_in, const FeatureParams& prm, FeatureOut& o) {
  const int N_SCANS = prm.n_scans;
  const double scanPeriod = 0.1;
  const int groundScanInd = 7;
  const double laderH = 0.56;
  const float Ground_scan_range[16] = {2.66f, 3.04f, 3.56f, 4.30f, 5.44f, 7.41f, 11.63f, 27.12f, 0, 0, 0, 0, 0, 0, 0, 0};

  // ---- removeClosedPointCloud (:732-763); NaN removal (:112) folded in ----
  std::vector<int> keep;
  keep.reserve(n_in);
  {
    const float th1 = (float)prm.minimum_range, th2 = (float)prm.maximum_range;
    for (int i = 0; i < n_in; i++) {
      const float x = in_xyzi[4 * (size_t)i], y = in_xyzi[4 * (size_t)i + 1], z = in_xyzi[4 * (size_t)i + 2];
      if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;
      float dis = x * x + y * y + z * z;
      if (dis < th1 * th1) continue;
      if (dis > th2 * th2) continue;
      if (x < 0 && (double)absf(y) < 0.5) continue;
      keep.push_back(i);
    }
  }
  int cloudSize = (int)keep.size();
  o = FeatureOut();
  if (cloudSize == 0) return;

  // ---- ring id + relTime (:116-213) ----
  auto P = [&](int kept, int c) { return in_xyzi[4 * (size_t)keep[kept] + c]; };
  float startOri = -std::atan2(P(0, 1), P(0, 0));
  float endOri = (float)(-std::atan2(P(cloudSize - 1, 1), P(cloudSize - 1, 0)) + 2 * M_PI);
  if (endOri - startOri > 3 * M_PI)
    endOri -= 2 * M_PI;
  else if (endOri - startOri < M_PI)
    endOri += 2 * M_PI;

  bool halfPassed = false;
  int count = cloudSize;
  std::vector<std::vector<float>> scans(N_SCANS);   // x,y,z,intensity
  std::vector<std::vector<int>> scans_src(N_SCANS), scans_int(N_SCANS);
  for (int i = 0; i < cloudSize; i++) {
    float px = P(i, 0), py = P(i, 1), pz = P(i, 2);
    int point_intensity = (int)P(i, 3);
    float verticalAngle = (float)(std::atan(pz / std::sqrt(px * px + py * py)) * 180 / M_PI);
    int scanID = 0;
    if (N_SCANS == 16) {
      scanID = int((verticalAngle + 15) / 2 + 0.5);
      if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
    } else if (N_SCANS == 32) {
      scanID = int((verticalAngle + 92.0 / 3.0) * 3.0 / 4.0);
      if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
    } else if (N_SCANS == 64) {
      if (verticalAngle >= -8.83)
        scanID = int((2 - verticalAngle) * 3.0 + 0.5);
      else
        scanID = N_SCANS / 2 + int((-8.83 - verticalAngle) * 2.0 + 0.5);
      if (verticalAngle > 2 || verticalAngle < -24.33 || scanID > 50 || scanID < 0) { count--; continue; }
    } else {
      return;
    }
    float ori = -std::atan2(py, px);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2)
        ori += 2 * M_PI;
      else if (ori > startOri + M_PI * 3 / 2)
        ori -= 2 * M_PI;
      if (ori - startOri > M_PI) halfPassed = true;
    } else {
      ori += 2 * M_PI;
      if (ori < endOri - M_PI * 3 / 2)
        ori += 2 * M_PI;
      else if (ori > endOri + M_PI / 2)
        ori -= 2 * M_PI;
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    float inten = (float)(scanID + scanPeriod * relTime);
    scans[scanID].insert(scans[scanID].end(), {px, py, pz, inten});
    scans_src[scanID].push_back(keep[i]);
    scans_int[scanID].push_back(point_intensity);
  }
  cloudSize = count;
  o.cloud_size = cloudSize;

  // ---- concatenate rings (:217-231) ----
  std::vector<float>& cloud = o.cloud;
  std::vector<int> intensity_num;
  o.scan_start.assign(N_SCANS, 0);
  o.scan_end.assign(N_SCANS, 0);
  for (int i = 0; i < N_SCANS; i++) {
    o.scan_start[i] = (int)(cloud.size() / 4) + 5;
    cloud.insert(cloud.end(), scans[i].begin(), scans[i].end());
    o.src_index.insert(o.src_index.end(), scans_src[i].begin(), scans_src[i].end());
    intensity_num.insert(intensity_num.end(), scans_int[i].begin(), scans_int[i].end());
    o.scan_end[i] = (int)(cloud.size() / 4) - 5;
  }
  std::vector<int> intensity_num2 = intensity_num;
  const int NP = cloudSize + 8;  // head-room: the occlusion pass writes index cloudSize (:454)
  auto X = [&](int i) { return cloud[4 * (size_t)i]; };
  auto Y = [&](int i) { return cloud[4 * (size_t)i + 1]; };
  auto Z = [&](int i) { return cloud[4 * (size_t)i + 2]; };

  o.range_vec.assign(NP, 0.f);
  o.scan_angle.assign(NP, 0.f);
  o.curvature.assign(NP, 0.f);
  o.inten_curvature.assign(NP, 0.f);
  o.curvature2.assign(NP, 0.f);
  o.distance_source.assign(NP, 0.f);
  o.other_source.assign(NP, 0.f);
  o.neighbor_picked.assign(NP, 0);
  o.inten_neighbor_picked.assign(NP, 0);
  o.label.assign(NP, 0);
  o.inten_label.assign(NP, 0);
  o.ground_marked.assign(NP, 0);
  std::vector<float>& range_vec = o.range_vec;
  std::vector<float>& scan_angle = o.scan_angle;

  // ---- range (:234-237) ----
  for (int i = 0; i < cloudSize; i++) range_vec[i] = std::sqrt(X(i) * X(i) + Y(i) * Y(i) + Z(i) * Z(i));

  // ---- incidence angle for near points (:239-255), double arithmetic, sequential sums ----
  for (int i = 5; i < cloudSize - 5; i++) {
    if (range_vec[i] < 2) {
      double a[3] = {X(i + 5), Y(i + 5), Z(i + 5)}, b[3] = {X(i - 5), Y(i - 5), Z(i - 5)}, now[3] = {X(i), Y(i), Z(i)};
      double c[3], ab[3], nc[3];
      for (int d = 0; d < 3; d++) {
        c[d] = (a[d] + b[d]) / 2;
        ab[d] = a[d] - b[d];
        nc[d] = now[d] - c[d];
      }
      double nrm[3] = {ab[1] * nc[2] - ab[2] * nc[1], ab[2] * nc[0] - ab[0] * nc[2], ab[0] * nc[1] - ab[1] * nc[0]};
      double dot = (nrm[0] * now[0] + nrm[1] * now[1]) + nrm[2] * now[2];
      double n1 = std::sqrt((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
      double n2 = std::sqrt((now[0] * now[0] + now[1] * now[1]) + now[2] * now[2]);
      scan_angle[i] = (float)(dot / (n1 * n2));
      if (scan_angle[i] < 0) scan_angle[i] = -scan_angle[i];
    }
  }
  // ---- near-range intensity smoothing with int truncation after every += (:257-268) ----
  for (int i = 5; i < cloudSize - 5; i++) {
    if (scan_angle[i] < 0.07 && range_vec[i] < 2) {
      intensity_num[i] = (int)(0.9 * intensity_num2[i]);
      for (int j = -5; j < 6; j++)
        if (j != 0) intensity_num[i] = (int)(intensity_num[i] + 0.005 * intensity_num2[i + j]);
    }
  }
  // ---- curvatures (:270-306) ----
  for (int i = 5; i < cloudSize - 5; i++) {
    float diffX = X(i - 5) + X(i - 4) + X(i - 3) + X(i - 2) + X(i - 1) - 10 * X(i) + X(i + 1) + X(i + 2) + X(i + 3) + X(i + 4) + X(i + 5);
    float diffY = Y(i - 5) + Y(i - 4) + Y(i - 3) + Y(i - 2) + Y(i - 1) - 10 * Y(i) + Y(i + 1) + Y(i + 2) + Y(i + 3) + Y(i + 4) + Y(i + 5);
    float diffZ = Z(i - 5) + Z(i - 4) + Z(i - 3) + Z(i - 2) + Z(i - 1) - 10 * Z(i) + Z(i + 1) + Z(i + 2) + Z(i + 3) + Z(i + 4) + Z(i + 5);
    const int* q = &intensity_num[i];
    float diffI = (float)(q[-5] + q[-4] + q[-3] + q[-2] + q[-1] - 10 * q[0] + q[1] + q[2] + q[3] + q[4] + q[5]);
    float dis_factor = (float)(2.0 / (1.0 + range_vec[i] / 20.0));
    if (dis_factor < 0.2) dis_factor = (float)0.2;
    o.curvature[i] = (diffX * diffX + diffY * diffY + diffZ * diffZ) * dis_factor;
    o.distance_source[i] = (float)(0.5 + dis_factor);
    float inten_factor = 1;
    if (scan_angle[i] < 0.07 && range_vec[i] < 2) {
      inten_factor = (float)(scan_angle[i] * 10 + 0.6);
      o.inten_curvature[i] = (float)((scan_angle[i] + 0.3) * diffI);
    } else {
      inten_factor = 3;
      o.inten_curvature[i] = diffI;
    }
    o.other_source[i] = inten_factor;
    const float* r = &range_vec[i];
    float diff_range = (float)(r[-5] + r[-4] + r[-3] + r[-2] + r[-1] - 10.0 * r[0] + r[1] + r[2] + r[3] + r[4] + r[5]);
    o.curvature2[i] = absf(diff_range * dis_factor);
  }

  // ---- ground marking + weighted plane (:307-431) ----
  std::vector<double> nearGround;  // xyz
  std::vector<double> laserweight;
  double center[3] = {0, 0, 0};
  double groundweights = 0;
  int groundsize = 0;
  size_t scanStart_ind = 0;
  for (int i = 0; i < groundScanInd && i < N_SCANS; i++) {
    const size_t ring_n = scans[i].size() / 4;
    // `col_ind < size() - 5` is size_t arithmetic in the reference and wraps for size < 5
    // (SURVEY App. A.15); generators guarantee >= 11 points, we clamp instead of wrapping.
    for (size_t col_ind = 5; ring_n >= 5 && col_ind < ring_n - 5; col_ind++) {
      int cloudInd = (int)(scanStart_ind + col_ind);
      float diff_range_th = (float)(0.8 * (1.0 + i / (groundScanInd - 1)));
      float diff_range = absf(range_vec[cloudInd] - Ground_scan_range[i]);
      double groundweight = 1.5 - i / (groundScanInd - 1);
      if (diff_range < diff_range_th) {
        if (scans[i][4 * col_ind + 2] < 0.3) {
          o.ground_marked[cloudInd] = 1;
          for (int n = -5; n < 5; n++) {
            if (absf(range_vec[cloudInd + n] - range_vec[cloudInd]) < diff_range_th / 2) {
              o.ground_marked[cloudInd + n] = 1;
              o.ground_points.push_back(cloudInd + n);
              double tmp[3] = {X(cloudInd + n), Y(cloudInd + n), Z(cloudInd + n)};
              for (int d = 0; d < 3; d++) center[d] = center[d] + groundweight * tmp[d];
              groundweights = groundweights + groundweight;
              nearGround.insert(nearGround.end(), tmp, tmp + 3);
              laserweight.push_back(groundweight);
              groundsize = groundsize + 1;
            }
          }
        }
      }
    }
    scanStart_ind += ring_n;
  }
  o.ground_size = groundsize;
  if (groundsize != 0) {
    for (int d = 0; d < 3; d++) center[d] = center[d] / groundweights;
    double covMat[9] = {0};
    for (int j = 0; j < groundsize; j++) {
      double t[3] = {nearGround[3 * j] - center[0], nearGround[3 * j + 1] - center[1], nearGround[3 * j + 2] - center[2]};
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) covMat[a * 3 + b] = covMat[a * 3 + b] + laserweight[j] * t[a] * t[b];
    }
    for (int a = 0; a < 9; a++) covMat[a] = covMat[a] / groundweights;
    double evals[3], V[9];
    eigh3(covMat, evals, V);
    for (int d = 0; d < 3; d++) o.ground_evals[d] = evals[d];
    double n[3] = {V[0], V[3], V[6]};
    double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int d = 0; d < 3; d++) n[d] /= nn;
    if (center[0] * n[0] + center[1] * n[1] + center[2] * n[2] < 0)
      for (int d = 0; d < 3; d++) n[d] = -n[d];
    double distance = 0, groundsource1 = 0;
    for (int j = 0; j < groundsize; j++) {
      const double* p = &nearGround[3 * j];
      double t[3] = {p[0] - center[0], p[1] - center[1], p[2] - center[2]};
      double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
      // Eigen normalized(): divides by the norm when it is > 0, else leaves the vector.
      if (tn > 0)
        for (int d = 0; d < 3; d++) t[d] /= tn;
      double distanceweight = 1 - 100 * std::fabs(n[0] * t[0] + n[1] * t[1] + n[2] * t[2]);
      if (distanceweight < 0) distanceweight = 0.1;
      groundsource1 += distanceweight;
      distance += distanceweight * (n[0] * p[0] + n[1] * p[1] + n[2] * p[2]);
    }
    distance = distance / groundsource1;
    groundsource1 = groundsource1 / groundsize;
    if ((distance / laderH) > 1.1 || (distance / laderH) < 0.9) distance = laderH;
    if (groundsource1 < 0.9) distance = 0.9 * laderH + 0.1 * distance;
    o.groundparam[0] = n[0]; o.groundparam[1] = n[1]; o.groundparam[2] = n[2];
    o.groundparam[3] = V[1]; o.groundparam[4] = V[4]; o.groundparam[5] = V[7];
    o.groundparam[6] = V[2]; o.groundparam[7] = V[5]; o.groundparam[8] = V[8];
    o.groundparam[9] = distance;
    o.groundparam[10] = 1 - groundsource1;
  }

  // ---- occlusion / parallel-beam masking (:433-456) ----
  for (int i = 5; i < cloudSize - 5; i++) {
    float depth1 = range_vec[i], depth2 = range_vec[i + 1];
    if (depth1 - depth2 > 0.04 * depth2) {
      for (int l = -5; l <= 0; l++) o.neighbor_picked[i + l] = 1;
    } else if (depth2 - depth1 > 0.04 * depth1) {
      for (int l = 1; l <= 6; l++) o.neighbor_picked[i + l] = 1;
    }
  }

  // ---- per ring x 6 sextants: sort + greedy selection (:469-644) ----
  std::vector<int> cloudSortInd(NP), intenSortInd(NP);
  for (int i = 0; i < NP; i++) cloudSortInd[i] = intenSortInd[i] = i;
  auto gap2 = [&](int a, int b) {
    float dx = X(a) - X(b), dy = Y(a) - Y(b), dz = Z(a) - Z(b);
    return dx * dx + dy * dy + dz * dz;
  };
  for (int i = 0; i < N_SCANS; i++) {
    if (o.scan_end[i] - o.scan_start[i] < 10) continue;
    for (int j = 0; j < 6; j++) {
      int sp = o.scan_start[i] + (o.scan_end[i] - o.scan_start[i]) * j / 6;
      int ep = o.scan_start[i] + (o.scan_end[i] - o.scan_start[i]) * (j + 1) / 6 - 1;
      // std::sort with comp(i,j)=curv[i]<curv[j] is unstable; tie policy (ours): (key, index).
      std::sort(cloudSortInd.begin() + sp, cloudSortInd.begin() + ep + 1, [&](int a, int b) {
        return o.curvature[a] < o.curvature[b] || (o.curvature[a] == o.curvature[b] && a < b);
      });
      std::sort(intenSortInd.begin() + sp, intenSortInd.begin() + ep + 1, [&](int a, int b) {
        return o.inten_curvature[a] < o.inten_curvature[b] || (o.inten_curvature[a] == o.inten_curvature[b] && a < b);
      });

      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = cloudSortInd[k];
        if (o.neighbor_picked[ind] == 0 && o.ground_marked[ind] != 1 && o.curvature[ind] > 0.1 && o.curvature2[ind] > 0.3) {
          largestPickedNum++;
          if (largestPickedNum <= 20) {
            o.label[ind] = 2;
            o.corner_sharp.push_back(ind);
            o.corner_sharp_w.push_back(o.distance_source[ind] + 1);
            o.corner_less_sharp.push_back(ind);
          } else if (largestPickedNum <= 21) {
            o.label[ind] = 1;
            o.corner_less_sharp.push_back(ind);
          } else {
            break;
          }
          o.neighbor_picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if (gap2(ind + l, ind + l - 1) > 0.05) break;
            o.neighbor_picked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if (gap2(ind + l, ind + l + 1) > 0.05) break;
            o.neighbor_picked[ind + l] = 1;
          }
        }
      }

      int smallestPickedNum = 0;
      for (int k = sp; k <= ep; k++) {
        int ind = cloudSortInd[k];
        if (o.neighbor_picked[ind] == 0 && o.curvature[ind] < 0.3 && o.curvature2[ind] < 0.4) {
          smallestPickedNum++;
          if (smallestPickedNum <= 40) {
            o.label[ind] = -1;
            o.surf_flat.push_back(ind);
            o.surf_flat_w.push_back(o.distance_source[ind]);
          } else {
            break;
          }
          o.neighbor_picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if (gap2(ind + l, ind + l - 1) > 0.05) break;
            o.neighbor_picked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if (gap2(ind + l, ind + l + 1) > 0.05) break;
            o.neighbor_picked[ind + l] = 1;
          }
        }
      }

      for (int k = sp; k <= ep; k++)
        if (o.label[k] <= 0) o.surf_less_flat.push_back(k);

      int largestPickedNum2 = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = intenSortInd[k];
        if (o.inten_neighbor_picked[ind] == 0 && o.ground_marked[ind] != 1 && o.inten_curvature[ind] > 65 && o.label[ind] != 2 && o.label[ind] != 1) {
          largestPickedNum2++;
          if (largestPickedNum2 <= 20) {
            o.inten_label[ind] = 2;
            o.inten_sharp.push_back(ind);
            o.inten_sharp_w.push_back(o.other_source[ind]);
            o.inten_less_sharp.push_back(ind);
          } else if (largestPickedNum2 <= 21) {
            o.inten_label[ind] = 1;
            o.inten_less_sharp.push_back(ind);
            o.corner_less_sharp.push_back(ind);
          } else {
            break;
          }
          o.inten_neighbor_picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            float diffI = (float)(intensity_num[ind + l] - intensity_num[ind + l - 1]);
            if (absf(diffI) > 35) break;
            o.inten_neighbor_picked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            float diffI = (float)(intensity_num[ind + l] - intensity_num[ind + l + 1]);
            if (absf(diffI) > 35) break;
            o.inten_neighbor_picked[ind + l] = 1;
          }
        }
      }
    }
  }
  // ---- intensity-edge merge (:645-663) ----
  if (prm.use_intensity) {
    double sharp = (double)o.corner_sharp.size(), plane = (double)o.surf_flat.size();
    double plane_sharp = sharp / plane;
    if (plane_sharp < 0.3) o.inten_merged = 1;
  }
  o.intensity_num = intensity_num;
}

}  // namespace orc
