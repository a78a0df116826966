// gslam_b200/plugin/dataset_synth.cpp -> libgslamDB_synth.so
//
// A synthetic GSLAM::Dataset (GSLAM/core/Dataset.h:100-162) for files with the extension ".synth", registered with
// GSLAM_REGISTER_DATASET (GSLAM/core/GSLAM.h:35-41) and found by GSLAM::Dataset::open() through
// Registry::load("gslamDB_synth") (Dataset.h:141-158) -- so that `gslam play b200_features metric_time -dataset x.synth` and the
// reference's player (GSLAM/plugins/play/main.cpp:15,126-132: grabFrame() -> publish on "dataset/frame") run unmodified on a
// stream whose bytes are known: the frames are EXACTLY gslam_b200/synth.py::synth_stream (same integer-only generator, restated in
// C++; tests/test_plugins.py::test_synth_dataset_frames_equal_python_generator compares them bit for bit).
//
// x.synth is a small text file of "key value" lines (all optional):
//     width 1920 / height 1080 / frames 30 / seed 7 / stereo 0 / disparity 14 / fx fy cx cy (pinhole, default 718-ish KITTI-like)
// Frames carry the GSLAM frame API a feature / SLAM plugin needs: getImage, getCamera, setKeyPoints / getKeyPoints / getDescriptor
// (GSLAM/core/Map.h:283-321).  stereo 1 delivers two images per frame (camera 1 = the same scene displaced by `disparity` px).
#include <GSLAM/core/GSLAM.h>
#include <GSLAM/core/Dataset.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

namespace {

// ---- the integer frame generator of gslam_b200/synth.py, restated -------------------------------------------------------------
inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t rand_u64(int64_t seed, int64_t stream, uint64_t i) {
  // Python: base = (seed * 0x100000001B3 + stream * 0x9E3779B1 + 0x1234567) & (2^64 - 1) with unbounded ints; seeds are small and
  // non-negative here, so the wrap-around product is the same thing
  const uint64_t base = (uint64_t)seed * 0x100000001B3ull + (uint64_t)stream * 0x9E3779B1ull + 0x1234567ull;
  return splitmix64(base + i * 0xD1342543DE82EF95ull);
}

void synth_scene(int W, int H, int64_t seed, std::vector<uint8_t>& img) {
  img.assign((size_t)W * H, 64);
  const int n_shapes = std::max(8, (int)(3000.0 * ((double)W * H) / 921600.0));
  for (int i = 0; i < n_shapes; ++i) {
    uint64_t r[6];
    for (int k = 0; k < 6; ++k) r[k] = rand_u64(seed, 1, (uint64_t)i * 6 + k);
    const int64_t cx = (int64_t)(r[0] % (uint64_t)W), cy = (int64_t)(r[1] % (uint64_t)H);
    const int64_t sw = (int64_t)(r[2] % 37) + 4, sh = (int64_t)(r[3] % 37) + 4;
    const uint8_t gray = (uint8_t)(r[4] % 256);
    const int kind = (int)(r[5] % 2);
    const int64_t x0 = std::max<int64_t>(0, cx - sw / 2), x1 = std::min<int64_t>(W, cx + sw / 2 + 1);
    const int64_t y0 = std::max<int64_t>(0, cy - sh / 2), y1 = std::min<int64_t>(H, cy + sh / 2 + 1);
    if (x1 <= x0 || y1 <= y0) continue;
    const int64_t rad = sw / 2;
    for (int64_t y = y0; y < y1; ++y)
      for (int64_t x = x0; x < x1; ++x)
        if (kind == 0 || (x - cx) * (x - cx) + (y - cy) * (y - cy) <= rad * rad) img[(size_t)y * W + x] = gray;
  }
}

// frame k of synth_stream(width, height, n_frames, seed); `shift_x` displaces the crop (stereo: the right eye)
void synth_frame(const std::vector<uint8_t>& big, int bigW, int width, int height, int64_t seed, int k, int shift_x, uint8_t* out) {
  const int64_t nseed = seed * 1000003 + k;
  const uint64_t plane = (uint64_t)width * height;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const uint64_t idx = (uint64_t)y * width + x;
      int64_t acc = 0;
      for (int half = 0; half < 2; ++half) {
        const uint64_t v = rand_u64(nseed, 2, half * plane + idx);
        for (int b = 0; b < 6; ++b) acc += (int64_t)((v >> (8 * b)) % 6);
      }
      const int64_t a = acc - 30;
      const int64_t noise = (a >= 0) ? a / 2 : -((-a + 1) / 2);  // Python floor division
      const int64_t px = (int64_t)big[(size_t)(5 * k + y) * bigW + 3 * k + x + shift_x] + noise;
      out[idx] = (uint8_t)std::min<int64_t>(255, std::max<int64_t>(0, px));
    }
}

class SynthFrame : public GSLAM::MapFrame {
 public:
  SynthFrame(GSLAM::FrameID id, double t, const GSLAM::GImage& left, const GSLAM::GImage& right, const GSLAM::Camera& cam, double baseline)
      : GSLAM::MapFrame(id, t), _left(left), _right(right), _cam(cam), _baseline(baseline) {}
  std::string type() const override { return _right.empty() ? "FrameMono" : "FrameStereo"; }
  int cameraNum() const override { return _right.empty() ? 1 : 2; }
  int imageChannels(int idx = 0) const override { return GSLAM::IMAGE_GRAY; }
  GSLAM::GImage getImage(int idx = 0, int channalMask = GSLAM::IMAGE_UNDEFINED) override { return idx == 1 ? _right : _left; }
  GSLAM::Camera getCamera(int idx = 0) override { return _cam; }
  GSLAM::SE3 getCameraPose(int idx = 0) const override {
    return idx == 1 ? GSLAM::SE3(GSLAM::SO3(), GSLAM::Point3d(_baseline, 0, 0)) : GSLAM::SE3();
  }
  // the feature carrier a tracking plugin fills (Map.h:310-321)
  int keyPointNum() const override { return (int)_kps.size(); }
  bool setKeyPoints(const std::vector<GSLAM::KeyPoint>& keypoints, const GSLAM::GImage& descriptors = GSLAM::GImage()) override {
    _kps = keypoints; _desc = descriptors; return true;
  }
  bool getKeyPoint(int idx, GSLAM::Point2f& pt) const override {
    if (idx < 0 || idx >= (int)_kps.size()) return false;
    pt = _kps[idx].pt; return true;
  }
  bool getKeyPoint(int idx, GSLAM::KeyPoint& pt) const override {
    if (idx < 0 || idx >= (int)_kps.size()) return false;
    pt = _kps[idx]; return true;
  }
  bool getKeyPoints(std::vector<GSLAM::KeyPoint>& keypoints) const override { keypoints = _kps; return true; }
  GSLAM::GImage getDescriptor(int idx = -1) const override {
    if (idx < 0) return _desc;
    if (idx >= _desc.rows) return GSLAM::GImage();
    return GSLAM::GImage(1, _desc.cols, _desc.type(), _desc.data + (size_t)idx * _desc.cols * _desc.elemSize(), true);
  }

 private:
  GSLAM::GImage _left, _right, _desc;
  GSLAM::Camera _cam;
  double _baseline;
  std::vector<GSLAM::KeyPoint> _kps;
};

class DatasetSynth : public GSLAM::Dataset {
 public:
  DatasetSynth() : _w(1920), _h(1080), _n(30), _seed(7), _stereo(0), _disp(14), _k(0), _opened(false), _fx(718.856), _fy(718.856), _cx(-1), _cy(-1) {}
  std::string type() const override { return "DatasetSynth"; }
  bool isOpened() override { return _opened; }
  bool isLive() const override { return false; }

  bool open(const std::string& path) override {
    std::ifstream f(path.c_str());
    std::string key;
    double v;
    while (f >> key >> v) {
      if (key == "width") _w = (int)v; else if (key == "height") _h = (int)v; else if (key == "frames") _n = (int)v;
      else if (key == "seed") _seed = (int64_t)v; else if (key == "stereo") _stereo = (int)v; else if (key == "disparity") _disp = (int)v;
      else if (key == "fx") _fx = v; else if (key == "fy") _fy = v; else if (key == "cx") _cx = v; else if (key == "cy") _cy = v;
    }
    if (_w < 64 || _h < 64 || _w > 16384 || _h > 16384 || _n < 1 || _n > 100000 || _seed < 0 || _disp < 0 || _disp > 512) {
      LOG(ERROR) << "DatasetSynth: unreasonable parameters in " << path;
      return false;
    }
    if (_cx < 0) _cx = 0.5 * _w;
    if (_cy < 0) _cy = 0.5 * _h;
    // the scene the whole stream crops from: (width + 3 n) x (height + 5 n), plus the stereo displacement
    _bigW = _w + 3 * _n + (_stereo ? _disp : 0);
    // NB the Python generator builds its scene for width + 3 n exactly; the stereo margin is appended on the right only when
    // stereo is on (a mono stream stays bit-identical to synth.synth_stream)
    synth_scene(_bigW, _h + 5 * _n, _seed, _big);
    _camera = GSLAM::Camera({(double)_w, (double)_h, _fx, _fy, _cx, _cy});
    _k = 0;
    _opened = true;
    return true;
  }

  GSLAM::FramePtr grabFrame() override {
    if (!_opened || _k >= _n) return GSLAM::FramePtr();
    GSLAM::GImage left(_h, _w, GSLAM::GImageType<uchar, 1>::Type), right;
    synth_frame(_big, _bigW, _w, _h, _seed, _k, 0, left.data);
    if (_stereo) {
      right = GSLAM::GImage(_h, _w, GSLAM::GImageType<uchar, 1>::Type);
      synth_frame(_big, _bigW, _w, _h, _seed, _k, _disp, right.data);  // the right eye sees the scene displaced by the disparity
    }
    const double baseline = _stereo ? 0.11 : 0.0;
    GSLAM::FramePtr fr(new SynthFrame(_k + 1, 0.05 * _k, left, right, _camera, baseline));
    ++_k;
    return fr;
  }

 private:
  int _w, _h, _n;
  int64_t _seed;
  int _stereo, _disp, _k, _bigW;
  bool _opened;
  double _fx, _fy, _cx, _cy;
  std::vector<uint8_t> _big;
  GSLAM::Camera _camera;
};

}  // namespace

GSLAM_REGISTER_DATASET(DatasetSynth, synth)
