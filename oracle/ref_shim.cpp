// oracle/ref_shim.cpp — NOT a restatement: a thin extern "C" window onto the UNMODIFIED reference headers, compiled from
// where they lie under /root/reference into oracle/_ref/libgslam_ref.so (recipe: oracle/Makefile, target `ref`).
// TEST INFRASTRUCTURE ONLY.  It pins the parts of the hot path that DO exist in the reference tree:
//   * Vocabulary::DistanceFactory::hamming32          GSLAM/core/Vocabulary.h:485-491
//   * SE3 inverse / point transform / product / exp / log   GSLAM/core/SE3.h:100-131,205-287  (pose conventions of BA)
//   * sizeof / layout of the carrier PODs             Map.h:122-195, Optimizer.h:106-172, SE3.h:337-339, SIM3.h:290-291
//   * Undistorter::undistort + the remap table of prepareReMap   GSLAM/core/Undistorter.h:120-348 (the frame-undistortion row)
// No reference source is copied into this repository; the .so is git-ignored and travels to the GPU box prebuilt.
#include <cstring>
#include <GSLAM/core/GSLAM.h>
#include <GSLAM/core/Optimizer.h>
#include <GSLAM/core/Vocabulary.h>
#include <GSLAM/core/Undistorter.h>
#include <sstream>

using namespace GSLAM;

static SE3 mk(const double* p) { return SE3(SO3(p[0], p[1], p[2], p[3]), Point3d(p[4], p[5], p[6])); }
static void put(const SE3& T, double* o) {
  const SO3& r = T.get_rotation();
  const Point3d& t = T.get_translation();
  o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; o[4] = t.x; o[5] = t.y; o[6] = t.z;
}

extern "C" {
__attribute__((visibility("default"))) float ref_hamming32(const unsigned char* a, const unsigned char* b) {
  return Vocabulary::DistanceFactory::hamming32(a, b);
}
__attribute__((visibility("default"))) void ref_se3_inverse(const double* pose7, double* out7) { put(mk(pose7).inverse(), out7); }
__attribute__((visibility("default"))) void ref_se3_transform(const double* pose7, const double* p3, double* out3) {
  Point3d q = mk(pose7) * Point3d(p3[0], p3[1], p3[2]);
  out3[0] = q.x; out3[1] = q.y; out3[2] = q.z;
}
__attribute__((visibility("default"))) void ref_se3_mul(const double* a7, const double* b7, double* out7) { put(mk(a7) * mk(b7), out7); }
__attribute__((visibility("default"))) void ref_se3_exp(const double* d6, double* out7) {
  Vector<double, 6> v;
  for (int i = 0; i < 6; ++i) v[i] = d6[i];
  put(SE3::exp(v), out7);
}
__attribute__((visibility("default"))) void ref_se3_log(const double* pose7, double* out6) {
  Vector<double, 6> v = mk(pose7).log();
  for (int i = 0; i < 6; ++i) out6[i] = v[i];
}
// The in-memory layout the BA boundary relies on: SIM3 = {SO3{x,y,z,w}, Point3d, scale} as 8 contiguous doubles.
__attribute__((visibility("default"))) void ref_sim3_raw(const double* pose7, double scale, double* out8) {
  SIM3 s(mk(pose7), scale);
  static_assert(sizeof(SIM3) == 64, "SIM3 layout");
  std::memcpy(out8, &s, 64);
}
__attribute__((visibility("default"))) int ref_sizeof(const char* name) {
  if (!std::strcmp(name, "KeyPoint")) return sizeof(KeyPoint);
  if (!std::strcmp(name, "SE3")) return sizeof(SE3);
  if (!std::strcmp(name, "SIM3")) return sizeof(SIM3);
  if (!std::strcmp(name, "Point3d")) return sizeof(Point3d);
  if (!std::strcmp(name, "BundleEdge")) return sizeof(BundleEdge);
  if (!std::strcmp(name, "KeyFrameEstimzation")) return sizeof(KeyFrameEstimzation);
  if (!std::strcmp(name, "MapPointEstimation")) return sizeof(MapPointEstimation);
  if (!std::strcmp(name, "GImage")) return sizeof(GImage);
  return -1;
}
// field offsets of KeyPoint, to prove gb_keypoint is field-compatible
__attribute__((visibility("default"))) int ref_keypoint_offsets(int* o7) {
  KeyPoint k;
  char* b = (char*)&k;
  o7[0] = (char*)&k.pt.x - b; o7[1] = (char*)&k.pt.y - b; o7[2] = (char*)&k.size - b; o7[3] = (char*)&k.angle - b;
  o7[4] = (char*)&k.response - b; o7[5] = (char*)&k.octave - b; o7[6] = (char*)&k.class_id - b;
  return 7;
}

// The reference's undistorter, run as is: cameras from parameter vectors (Camera.h:435-444), tables from prepareReMap, output from
// undistort().  tables: idx4 / coef4 / remap_x may be NULL; out may be NULL.  Returns 0 when both cameras are valid.
__attribute__((visibility("default"))) int ref_undistort(const double* cam_in, int n_in, const double* cam_out, int n_out, const unsigned char* img,
                                                         int channels, int* idx4, float* coef4, float* remap_x, unsigned char* out) {
  Camera in(std::vector<double>(cam_in, cam_in + n_in)), outc(std::vector<double>(cam_out, cam_out + n_out));
  if (!in.isValid() || !outc.isValid()) return 1;
  UndistorterImpl u(in, outc);  // (prepareReMap prints the two camera models on stdout)
  if (!u.valid) return 2;
  const size_t n = (size_t)outc.width() * outc.height();
  if (idx4) std::memcpy(idx4, u.remapIdx, n * 16);
  if (coef4) std::memcpy(coef4, u.remapCoef, n * 16);
  if (remap_x) std::memcpy(remap_x, u.remapX, n * 4);
  if (img && out) {
    // the reference reads up to one row + one pixel past the end of the image for taps of the last row: let it read zeros from a
    // padded buffer it does not own (GImage.h:169, copy = false) -- which is the value the product defines for those taps
    const size_t bytes = (size_t)in.width() * in.height() * channels;
    std::vector<unsigned char> padded(bytes + (size_t)(in.width() + 2) * channels, 0);
    std::memcpy(padded.data(), img, bytes);
    GImage src(in.height(), in.width(), channels == 1 ? GImageType<uchar, 1>::Type : GImageType<uchar, 3>::Type, padded.data(), false);
    GImage dst;
    if (!u.undistort(src, dst)) return 3;
    std::memcpy(out, dst.data, n * channels);
  }
  return 0;
}
}
