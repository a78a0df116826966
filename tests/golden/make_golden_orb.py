"""Golden fixtures of the ORB path, generated HERE from cv2 (OpenCV 4.13.0 wheel) — the CPU path GSLAM's SLAM plugins run.
Images come from gslam_b200.synth (integer-only generator); their sha256 is stored so a drifted generator is detected."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gslam_b200 import synth  # noqa: E402

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

CASES = [  # (name, w, h, nfeatures, seed, store_image)
    ("orb_320x240_n300", 320, 240, 300, 21, True),
    ("orb_480x360_n400", 480, 360, 400, 22, True),
    ("orb_752x480_n2000", 752, 480, 2000, 23, False),
    ("orb_1280x720_n1000", 1280, 720, 1000, 24, False),
    ("orb_1920x1080_n2000", 1920, 1080, 2000, 25, False),
]


def cv2_orb_canonical(img, nfeatures, **kw):
    """cv2.ORB_create(nfeatures).detectAndCompute -> records sorted by (octave, y_level, x_level)."""
    import cv2
    orb = cv2.ORB_create(nfeatures=nfeatures, **kw)
    kc, dc = orb.detectAndCompute(img, None)
    sf = np.float64(np.float32(kw.get("scaleFactor", 1.2)))
    rec = np.zeros(len(kc), KP_DTYPE)
    keys = []
    for i, k in enumerate(kc):
        rec[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
        s = np.float32(sf ** k.octave)
        inv = np.float32(1.0) / s
        keys.append((k.octave, int(np.rint(np.float32(k.pt[1]) * inv)), int(np.rint(np.float32(k.pt[0]) * inv)), i))
    keys.sort()
    order = [i for *_, i in keys]
    lvl = np.array([(o, y, x) for o, y, x, _ in keys], np.int32).reshape(-1, 3)
    if dc is None:
        dc = np.zeros((0, 32), np.uint8)
    return rec[order], dc[order], lvl


def make_orb():
    import cv2
    for name, w, h, n, seed, store in CASES:
        img = synth.synth_frame(w, h, seed)
        kps, desc, lvl = cv2_orb_canonical(img, n)
        out = dict(width=w, height=h, nfeatures=n, seed=seed, image_sha256=hashlib.sha256(img.tobytes()).hexdigest(),
                   kps=kps, desc=desc, level_yx=lvl, cv2_version=cv2.__version__)
        if store:
            out["image"] = img
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, len(kps))
    # stage fixtures on a small image
    img = synth.synth_frame(200, 150, 31)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    fd = cv2.FastFeatureDetector_create(20, True)
    f1 = fd.detect(img, None)
    fd0 = cv2.FastFeatureDetector_create(20, False)
    f0 = fd0.detect(img, None)
    rng = np.random.default_rng(0)
    ay = (rng.integers(-200000, 200000, 4000)).astype(np.float32); ax = (rng.integers(-200000, 200000, 4000)).astype(np.float32)
    ay[:8] = [0, 0, 1, -1, 1, -1, 5, 0]; ax[:8] = [0, 1, 0, 0, 1, -1, 5, -3]
    at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ay, ax)], np.float32)
    small = cv2.resize(img, (167, 125), interpolation=cv2.INTER_LINEAR_EXACT)
    np.savez_compressed(os.path.join(HERE, "orb_stages.npz"), image=img, gauss_kernel=k,
                        fast_nms=np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in f1], np.int32),
                        fast_all=np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in f0], np.int32),
                        atan_y=ay, atan_x=ax, atan=at, resized_167x125=small,
                        blur=cv2.sepFilter2D(img, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101))
    print("orb_stages.npz")


if __name__ == "__main__":
    make_orb()
