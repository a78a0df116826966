"""GPU parity of ORB extract (K1-K4) through the C-ABI: keypoints (u,v,octave,angle,response,size) and 256-bit
descriptors bit-exact against the cv2 golden fixtures and against the oracle on seeded frames."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import oracle
from gslam_b200 import capi, synth
from gslam_b200.api import Features

pytestmark = pytest.mark.gpu
GD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["orb_320x240_n300", "orb_480x360_n400", "orb_752x480_n2000", "orb_1280x720_n1000", "orb_1920x1080_n2000"]
FIELDS = ("octave", "x", "y", "size", "angle", "response", "class_id")


def assert_same(kps, desc, wk, wd):
    assert len(kps) == len(wk), (len(kps), len(wk))
    for f in FIELDS:
        assert np.array_equal(kps[f], wk[f]), f
    assert np.array_equal(desc, wd)


@pytest.mark.parametrize("name", CASES)
def test_golden_cv2(ctx, name):
    g = np.load(os.path.join(GD, name + ".npz"))
    img = g["image"] if "image" in g else synth.synth_frame(int(g["width"]), int(g["height"]), int(g["seed"]))
    if hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() != str(g["image_sha256"]):
        pytest.skip("synthetic image generator drifted from the fixture")
    kps, desc = ctx.orb_extract(img, int(g["nfeatures"]))
    assert_same(kps, desc, g["kps"], g["desc"])


@pytest.mark.parametrize("w,h,n,kw", [
    (517, 389, 350, {}),                                   # odd sizes, ragged tiles
    (517, 389, 200, dict(nlevels=5, fast_threshold=12)),
    (517, 389, 300, dict(scale_factor=1.35, nlevels=6)),
    (131, 97, 150, {}),                                    # upper levels too small to hold a keypoint
    (1000, 64, 100, {}),                                   # nothing fits the 31-px border vertically
    (2048, 1536, 5000, {}),                                # larger than BASELINE sizes
    (456, 477, 4000, dict(fast_threshold=30)),             # 477/1.2 lands on a half: cv2 sizes level 1 as 380x398 (not 397)
    (303, 249, 600, {}),                                   # same rule on both axes at different levels
])
def test_vs_oracle(ctx, w, h, n, kw):
    img = synth.synth_frame(w, h, 1000 + w)
    wk, wd = oracle.orb_extract(img, n, **kw)
    kps, desc = ctx.orb_extract(img, n, **kw)
    assert_same(kps, desc, wk, wd)


from _images import KIND_CASES, image_of_kind  # noqa: E402  (tests/ is on sys.path under pytest's rootdir conftest)


@pytest.mark.parametrize("kind,w,h,seed,n,nl,sf,ft", KIND_CASES)
def test_other_image_statistics_match_oracle(ctx, kind, w, h, seed, n, nl, sf, ft):
    """Image families unlike synth_frame -- white noise, binary noise, checkerboards (exact ties in FAST scores, Harris responses and
    moments everywhere), blurred noise, constant blocks, low contrast -- on which the oracle is pinned to live cv2
    (tests/test_oracle_orb.py::test_live_cv2_image_kinds): every keypoint field and descriptor bit against the oracle."""
    img = np.ascontiguousarray(image_of_kind(kind, w, h, seed))
    kw = dict(nlevels=nl, scale_factor=sf, fast_threshold=ft)
    wk, wd = oracle.orb_extract(img, n, **kw)
    kps, desc = ctx.orb_extract(img, n, **kw)
    assert_same(kps, desc, wk, wd)


def _dbg_level(ctx, level, cap):
    L = capi.lib()
    L.gb_dbg_orb_level.restype = C.c_int
    L.gb_dbg_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    w, h = C.c_int(), C.c_int()
    buf = np.zeros(cap, np.uint8)
    assert L.gb_dbg_orb_level(ctx.handle, level, buf.ctypes.data, cap, C.byref(w), C.byref(h)) == 0
    return buf[:w.value * h.value].reshape(h.value, w.value)


def _dbg_candidates(ctx, level):
    L = capi.lib()
    L.gb_dbg_orb_candidates.restype = C.c_int
    L.gb_dbg_orb_candidates.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    cap = 1 << 18
    pos = np.zeros(cap, np.uint32); sc = np.zeros(cap, np.uint8); n = C.c_int(); nk = C.c_int()
    assert L.gb_dbg_orb_candidates(ctx.handle, level, pos.ctypes.data, sc.ctypes.data, None, None, cap, C.byref(n), None, 0, C.byref(nk)) == 0
    return {(int(p & 0xffff), int(p >> 16)): int(s) for p, s in zip(pos[:n.value], sc[:n.value])}


@pytest.mark.parametrize("w,h", [(320, 240), (517, 389), (1280, 720)])
def test_pyramid_levels_match_oracle(ctx, w, h):
    """K1 alone: every pyramid level bit-exact (INTER_LINEAR_EXACT from the previous level)."""
    img = synth.synth_frame(w, h, 50 + w)
    ctx.orb_extract(img, 300)
    for l in range(8):
        assert np.array_equal(_dbg_level(ctx, l, w * h), oracle.orb_pyramid_level(img, l)), l


@pytest.mark.parametrize("w,h", [(320, 240), (517, 389), (1280, 720)])
def test_fast_candidates_match_oracle(ctx, w, h):
    """K2 alone: FAST-9/16 positions AND scores after NMS + border filter, every level (guards the ptxas VIMNMX3 miscompile)."""
    img = synth.synth_frame(w, h, 60 + w)
    ctx.orb_extract(img, 300)
    for l in range(8):
        lv = oracle.orb_pyramid_level(img, l)
        lh, lw = lv.shape
        xs, ys, sc = oracle.fast_detect(lv, 20, True)
        keep = (xs >= 31) & (xs < lw - 31) & (ys >= 31) & (ys < lh - 31)
        want = {(int(x), int(y)): int(s) for x, y, s in zip(xs[keep], ys[keep], sc[keep])}
        if lw <= 62 or lh <= 62:
            want = {}
        assert _dbg_candidates(ctx, l) == want, l


def test_stream_of_frames_same_context(ctx):
    frames = synth.synth_stream(640, 480, 4, seed=9)
    for f in frames:  # buffers are reused across frames: no state may leak from one frame to the next
        wk, wd = oracle.orb_extract(f, 500)
        kps, desc = ctx.orb_extract(f, 500)
        assert_same(kps, desc, wk, wd)


def test_flat_and_tiny_images(ctx):
    kps, desc = ctx.orb_extract(np.full((100, 120), 77, np.uint8), 100)
    assert len(kps) == 0 and desc.shape == (0, 32)
    kps, _ = ctx.orb_extract(synth.synth_frame(40, 40, 1), 100)
    assert len(kps) == 0


def test_high_contrast_extremes(ctx):
    rng = np.random.default_rng(3)
    img = (rng.integers(0, 2, (300, 400)) * 255).astype(np.uint8)  # binary noise: saturating blur, many ties in FAST score
    wk, wd = oracle.orb_extract(img, 400)
    kps, desc = ctx.orb_extract(img, 400)
    assert_same(kps, desc, wk, wd)
    img2 = np.zeros((300, 400), np.uint8); img2[::16, :] = 255; img2[:, ::16] = 255   # periodic grid: exact ties in Harris
    wk, wd = oracle.orb_extract(img2, 200)
    kps, desc = ctx.orb_extract(img2, 200, capacity=len(wk) + 16)
    assert_same(kps, desc, wk, wd)


def test_capacity_error_reports_required(ctx):
    img = synth.synth_frame(640, 480, 5)
    cfg = ctx.orb_cfg(nfeatures=500)
    kps = np.zeros(10, capi.KP_DTYPE); desc = np.zeros((10, 32), np.uint8); n = C.c_int(10)
    rc = ctx._lib.gb_orb_extract(ctx.handle, capi.ptr(img), 640, 480, C.byref(cfg), capi.ptr(kps), capi.ptr(desc), C.byref(n))
    assert rc == capi.GB_ERR_CAPACITY and n.value == 500
    assert not kps["x"].any()  # nothing written


def test_unsupported_config_rejected(ctx):
    img = synth.synth_frame(320, 240, 5)
    for kw in (dict(wta_k=3), dict(patch_size=25), dict(score_type=1), dict(first_level=1), dict(edge_threshold=10), dict(nlevels=0)):
        with pytest.raises(capi.GbError):
            ctx.orb_extract(img, 100, **kw)


def test_device_resident_extract_and_match(ctx):
    """extract -> match chained on the device; compared with oracle extract + oracle match."""
    frames = synth.synth_stream(1280, 720, 2, seed=4)
    cfg = ctx.orb_cfg(nfeatures=1000)
    f0, f1 = Features(ctx, 2304), Features(ctx, 2304)
    f0.extract(frames[0], 1280, 720, cfg)
    f1.extract(frames[1], 1280, 720, cfg)
    f1.match(f0)
    idx, d1, d2 = f1.matches()
    k0, de0 = oracle.orb_extract(frames[0], 1000); k1, de1 = oracle.orb_extract(frames[1], 1000)
    got1 = f1.download()
    assert_same(got1[0], got1[1], k1, de1)
    w = oracle.match_hamming(de1, de0)
    assert np.array_equal(idx, w[0]) and np.array_equal(d1, w[1]) and np.array_equal(d2, w[2])
    # the stream moves by (3,5) px: good matches must agree with that motion
    good = d1 < 40
    assert good.sum() > 200
    dx = k1["x"][good] - k0["x"][idx[good]]; dy = k1["y"][good] - k0["y"][idx[good]]
    assert np.median(np.abs(dx + 3)) < 1.5 and np.median(np.abs(dy + 5)) < 1.5
    f0.close(); f1.close()


@pytest.mark.parametrize("channels,rgb", [(3, False), (4, False), (3, True), (4, True)])
def test_colour_frames_are_converted_on_the_device(ctx, channels, rgb):
    """8UC3 / 8UC4 frames (what GSLAM's dataset plugins deliver, IO.h:86-110): gb_orb_extract_image == gray extraction of the
    cv2-rule gray image (oracle.to_gray), every field and every descriptor bit."""
    rng = np.random.default_rng(channels * 2 + rgb)
    base = synth.synth_frame(480, 360, seed=11).astype(np.int64)
    col = np.stack([np.clip(base + rng.integers(-40, 41, base.shape), 0, 255) for _ in range(channels)], -1).astype(np.uint8)
    kps, desc = ctx.orb_extract(col, 400, rgb=rgb)
    k0, d0 = ctx.orb_extract(oracle.to_gray(col, rgb=rgb), 400)
    assert len(kps) == len(k0) > 300 and np.array_equal(kps, k0) and np.array_equal(desc, d0)
