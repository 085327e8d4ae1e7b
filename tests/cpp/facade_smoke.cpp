// Reads two clouds (binary float32 x,y,z,1 records) and runs the reference call-site sequence
// (RGC_odometer.cpp:998-1015) through the C++ facade; prints the final transformation row-major.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../include/rgc/fast_gicp.hpp"

using PointT = rgc::PointXYZI;

static rgc::PointCloud<PointT>::Ptr load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  std::fseek(f, 0, SEEK_END);
  long bytes = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<float> raw(bytes / 4);
  if (std::fread(raw.data(), 4, raw.size(), f) != raw.size()) std::exit(2);
  std::fclose(f);
  auto cloud = std::make_shared<rgc::PointCloud<PointT>>();
  cloud->resize(raw.size() / 4);
  for (size_t i = 0; i < cloud->size(); i++) {
    (*cloud)[i].x = raw[4 * i]; (*cloud)[i].y = raw[4 * i + 1]; (*cloud)[i].z = raw[4 * i + 2];
    (*cloud)[i].intensity = (float)i;
  }
  return cloud;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  auto target = load(argv[1]);
  auto source = load(argv[2]);
  rgc::PointCloud<PointT> aligned;
  rgc::Matrix4f T2 = rgc::identity4();
  {
    rgc::FastGICP<PointT, PointT> vgicp;  // stack-local, constructed per frame like the reference
    vgicp.setMaximumIterations(25);
    vgicp.setMaxCorrespondenceDistance(2);
    vgicp.setTransformationEpsilon(1e-6);
    vgicp.setEuclideanFitnessEpsilon(1e-6);
    vgicp.setRANSACIterations(0);
    vgicp.setNumThreads(14);
    vgicp.setInputTarget(target);
    vgicp.setInputSource(source);
    vgicp.align(aligned, T2);
    double score = vgicp.getFitnessScore();
    const rgc::Matrix4f& T = vgicp.getFinalTransformation();
    std::printf("converged %d iterations %d fitness %.9g aligned %zu intensity_kept %d\n", (int)vgicp.hasConverged(), vgicp.lastResult().iterations,
                score, aligned.size(), (int)(aligned[7].intensity == 7.f));
    for (int r = 0; r < 4; r++) std::printf("%.9g %.9g %.9g %.9g\n", T[r], T[4 + r], T[8 + r], T[12 + r]);
    vgicp.setInputTarget(target);  // same shared_ptr: must be a no-op
  }
  {  // the exact sequence of RGC_odometer.cpp:998-1011, voxelised variant
    rgc::FastVGICP<PointT, PointT> vgicp;
    vgicp.setResolution(1);
    vgicp.setMaximumIterations(25);
    vgicp.setMaxCorrespondenceDistance(2);
    vgicp.setTransformationEpsilon(1e-6);
    vgicp.setNumThreads(14);
    vgicp.setInputTarget(target);
    vgicp.setInputSource(source);
    vgicp.align(aligned, T2);
    const rgc::Matrix4f& T = vgicp.getFinalTransformation();
    std::printf("vgicp converged %d iterations %d fitness %.9g\n", (int)vgicp.hasConverged(), vgicp.lastResult().iterations, vgicp.getFitnessScore());
    for (int r = 0; r < 4; r++) std::printf("%.9g %.9g %.9g %.9g\n", T[r], T[4 + r], T[8 + r], T[12 + r]);
  }
  {  // the frame's front end fused into setInput* (voxel filters of RGC_odometer.cpp:975-991)
    rgc::FastGICP<PointT, PointT> gicp;
    gicp.setMaxCorrespondenceDistance(2);
    const size_t nt = gicp.setInputTargetFiltered(target, 0.3f);
    const size_t ns = gicp.setInputSourceFiltered(source, 0.2f);
    gicp.align(aligned, T2);
    const rgc::Matrix4f& T = gicp.getFinalTransformation();
    std::printf("filtered converged %d iterations %d n_source %zu n_target %zu aligned %zu\n", (int)gicp.hasConverged(), gicp.lastResult().iterations, ns, nt,
                aligned.size());
    for (int r = 0; r < 4; r++) std::printf("%.9g %.9g %.9g %.9g\n", T[r], T[4 + r], T[8 + r], T[12 + r]);
  }
  {  // swapSourceAndTarget (fast_gicp_impl.hpp:49-57): the output cloud is the OLD target, fields included
    rgc::FastGICP<PointT, PointT> gicp;
    gicp.setMaxCorrespondenceDistance(2);
    gicp.setInputTarget(target);
    gicp.setInputSource(source);
    gicp.swapSourceAndTarget();
    gicp.align(aligned);
    std::printf("swap converged %d aligned %zu expect %zu intensity_kept %d\n", (int)gicp.hasConverged(), aligned.size(), target->size(),
                (int)(aligned.size() == target->size() && aligned[aligned.size() - 1].intensity == (float)(target->size() - 1)));
  }
  return 0;
}
