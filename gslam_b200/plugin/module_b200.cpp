// gslam_b200/plugin/module_b200.cpp -> libgslam_b200.so
//
// Svar module (GSLAM/core/Svar.h:71-76, loaded with GSLAM::Registry::load("b200"), Registry.h:48-77) that exposes the feature
// path the reference has no interface for (the ORB calls live in external SLAM plugins, README.md:130-133):
//   svar["gslam"]["b200"]["orb_extract"]   (GImage gray8, Svar cfg) -> {"keypoints": vector<KeyPoint>, "descriptors": GImage Nx32 8UC1}
//   svar["gslam"]["b200"]["match_hamming"] (GImage query Nx32, GImage train Mx32) -> {"trainIdx": vector<int>, "distance": vector<int>,
//                                                                                    "distance2": vector<int>}
//   svar["gslam"]["b200"]["extract_to_frame"] (FramePtr, Svar cfg) -> bool : getImage() -> orb_extract -> setKeyPoints()
//                                                                     (Map.h:287,311-312)
// Outputs are the reference's own carrier types: GSLAM::KeyPoint (Map.h:122-195) and an owning GImage (GImage.h:160-443).
// Functions do not throw; on failure they return an undefined Svar / false and log through GSLAM's LOG.
#include <GSLAM/core/GSLAM.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/gslam_b200.h"

namespace {

static_assert(sizeof(GSLAM::KeyPoint) == sizeof(gb_keypoint), "gb_keypoint must be field-compatible with GSLAM::KeyPoint");

struct SharedCtx {
  gb_ctx* ctx = nullptr;
  std::mutex mu;
  gb_ctx* get() {
    std::lock_guard<std::mutex> lk(mu);
    if (!ctx) {
      if (gb_ctx_create(svar.GetInt("b200.device", 0), &ctx) != GB_OK) {
        LOG(ERROR) << "gslam_b200: no usable CUDA device (" << gb_last_error(NULL) << "); there is no CPU fallback";
        ctx = nullptr;
      }
    }
    return ctx;
  }
};
SharedCtx& shared() {
  static SharedCtx s;
  return s;
}

gb_orb_cfg cfgFrom(GSLAM::Svar cfg) {
  gb_orb_cfg c;
  gb_orb_cfg_default(&c);
  if (cfg.isObject()) {
    c.nfeatures = cfg.get<int>("nfeatures", c.nfeatures);
    c.scale_factor = (float)cfg.get<double>("scaleFactor", c.scale_factor);
    c.nlevels = cfg.get<int>("nlevels", c.nlevels);
    c.edge_threshold = cfg.get<int>("edgeThreshold", c.edge_threshold);
    c.fast_threshold = cfg.get<int>("fastThreshold", c.fast_threshold);
  }
  return c;
}

bool extract(const GSLAM::GImage& img, GSLAM::Svar cfg, std::vector<GSLAM::KeyPoint>& kps, GSLAM::GImage& desc) {
  if (img.empty() || img.type() != GSLAM::GImageType<uchar, 1>::Type) {
    LOG(ERROR) << "gslam_b200 orb_extract: need a non-empty 8UC1 image (convert colour frames to gray first)";
    return false;
  }
  gb_ctx* ctx = shared().get();
  if (!ctx) return false;
  gb_orb_cfg c = cfgFrom(cfg);
  int n = 2 * c.nfeatures + 256;
  for (int attempt = 0; attempt < 2; ++attempt) {
    kps.resize(n);
    std::vector<uint8_t> d((size_t)n * 32);
    int got = n;
    // GImage has no row stride: row i at data + i*cols (GImage.h:378)
    const int rc = gb_orb_extract(ctx, img.data, img.cols, img.rows, &c, reinterpret_cast<gb_keypoint*>(kps.data()), d.data(), &got);
    if (rc == GB_ERR_CAPACITY && got > n) { n = got; continue; }
    if (rc != GB_OK) {
      LOG(ERROR) << "gslam_b200 orb_extract: " << gb_last_error(ctx);
      return false;
    }
    kps.resize(got);
    desc = GSLAM::GImage(got, 32, GSLAM::GImageType<uchar, 1>::Type, got ? d.data() : NULL, true);  // owning copy
    return true;
  }
  return false;
}

}  // namespace

EXPORT_SVAR_INSTANCE
REGISTER_SVAR_MODULE(b200) {
  GSLAM_REGISTER_GLOG_SINKS
  GSLAM_REGISTER_MESSENGER
  svar["gslam"]["b200"]["orb_extract"] = GSLAM::Svar::lambda([](GSLAM::GImage img, GSLAM::Svar cfg) -> GSLAM::Svar {
    std::vector<GSLAM::KeyPoint> kps;
    GSLAM::GImage desc;
    if (!extract(img, cfg, kps, desc)) return GSLAM::Svar();
    GSLAM::Svar out = GSLAM::Svar::object();
    out["keypoints"] = kps;
    out["descriptors"] = desc;
    return out;
  });
  svar["gslam"]["b200"]["match_hamming"] = GSLAM::Svar::lambda([](GSLAM::GImage q, GSLAM::GImage t) -> GSLAM::Svar {
    if (q.cols != 32 || (t.rows > 0 && t.cols != 32) || q.elemSize() != 1) {
      LOG(ERROR) << "gslam_b200 match_hamming: descriptors must be N x 32 8UC1";
      return GSLAM::Svar();
    }
    gb_ctx* ctx = shared().get();
    if (!ctx) return GSLAM::Svar();
    std::vector<int> idx(q.rows), d1(q.rows), d2(q.rows);
    if (gb_match_hamming(ctx, q.data, q.rows, t.data, t.rows, idx.data(), d1.data(), d2.data()) != GB_OK) {
      LOG(ERROR) << "gslam_b200 match_hamming: " << gb_last_error(ctx);
      return GSLAM::Svar();
    }
    GSLAM::Svar out = GSLAM::Svar::object();
    out["trainIdx"] = idx;
    out["distance"] = d1;
    out["distance2"] = d2;
    return out;
  });
  svar["gslam"]["b200"]["extract_to_frame"] = GSLAM::Svar::lambda([](GSLAM::FramePtr fr, GSLAM::Svar cfg) -> bool {
    if (!fr) return false;
    std::vector<GSLAM::KeyPoint> kps;
    GSLAM::GImage desc;
    if (!extract(fr->getImage(), cfg, kps, desc)) return false;
    return fr->setKeyPoints(kps, desc);  // Map.h:311-312
  });
}
