"""Frame undistortion on the device (gb_remap_*, csrc/remap.cu) against the reference's own output: the committed vectors written by
GSLAM::Undistorter (tests/golden/remap_opencv_96x72.npz) and, at the benchmark's frame size, the oracle restatement that
tests/test_oracle_remap.py pins to the reference byte for byte."""
import os

import numpy as np
import pytest

import oracle
from gslam_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def test_remap_equals_reference_vectors(ctx):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "remap_opencv_96x72.npz"))
    h, w = g["out1"].shape
    m = api.Remap(ctx, w, h, w, h, g["idx4"], g["coef4"], g["remap_x"])
    assert np.array_equal(m.apply(g["img"]), g["out1"])
    inside = (g["remap_x"] > 0).reshape(h, w)
    out3 = m.apply(g["rgb"])
    assert np.array_equal(out3[inside], g["out3"][inside]) and not out3[~inside].any()
    m.close()


def _bilinear_table(w_in, h_in, w_out, h_out, seed):
    """A smooth synthetic distortion in the reference's table format (indices / weights as prepareReMap lays them out)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h_out, 0:w_out].astype(np.float64)
    u = (x - w_out / 2) / w_out; v = (y - h_out / 2) / h_out
    r2 = u * u + v * v
    k = 1 + rng.uniform(-0.4, -0.1) * r2 + 0.1 * r2 * r2
    sx = (u * k * 1.15 + 0.5) * w_in; sy = (v * k * 1.15 + 0.5) * h_in
    out = (sx < 0) | (sy < 0) | (sx >= w_in) | (sy >= h_in)
    rx = np.where(out, -1, sx).astype(np.float32)
    ry = np.where(out, -1, sy).astype(np.float32)
    xi = rx.astype(np.int32); yi = ry.astype(np.int32)
    fx = rx - xi; fy = ry - yi; fxy = fx * fy
    idx = np.stack([yi * w_in + xi, yi * w_in + xi + 1, (yi + 1) * w_in + xi, (yi + 1) * w_in + xi + 1], -1).astype(np.int32)
    coef = np.stack([1 - fx - fy + fxy, fx - fxy, fy - fxy, fxy], -1).astype(np.float32)
    idx[out] = 0; coef[out] = 0
    return idx.reshape(-1, 4), coef.reshape(-1, 4), rx.reshape(-1)


@pytest.mark.parametrize("ch", [1, 3])
@pytest.mark.parametrize("size", [(1920, 1080, 1920, 1080), (752, 480, 640, 400), (97, 61, 33, 19)])
def test_remap_equals_oracle_at_frame_sizes(ctx, size, ch):
    w_in, h_in, w_out, h_out = size
    idx, coef, rx = _bilinear_table(w_in, h_in, w_out, h_out, seed=w_in)
    img = synth.synth_frame(w_in, h_in, seed=2)
    if ch == 3:
        img = np.stack([img, np.roll(img, 9, axis=1), 255 - img], axis=2)
    m = api.Remap(ctx, w_in, h_in, w_out, h_out, idx, coef, rx)
    got = m.apply(img)
    want = oracle.remap_apply(img, idx, coef, rx, (h_out, w_out))
    assert (rx >= 0).mean() > 0.5 and (idx.max() >= w_in * h_in or size[0] > 100)   # (last-row taps beyond the image occur)
    assert np.array_equal(got, want)
    again = m.apply(img)                                                          # idempotent on the cached staging buffers
    assert np.array_equal(again, got)
    m.close()


def test_remap_rejects_bad_arguments(ctx):
    idx, coef, rx = _bilinear_table(64, 48, 64, 48, 1)
    bad = idx.copy(); bad[5, 2] = -3
    with pytest.raises(api.GbError):
        api.Remap(ctx, 64, 48, 64, 48, bad, coef, rx)
    with pytest.raises(api.GbError):
        api.Remap(ctx, 0, 48, 64, 48, idx, coef, rx)
    m = api.Remap(ctx, 64, 48, 64, 48, idx, coef, rx)
    with pytest.raises(api.GbError):
        ctx._check(ctx._lib.gb_remap_apply(ctx._h, m._h, api.ptr(np.zeros((48, 64, 2), np.uint8)), 2, api.ptr(np.zeros((48, 64, 2), np.uint8))))
    m.close()
