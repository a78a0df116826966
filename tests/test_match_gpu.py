"""GPU parity of the 256-bit Hamming matcher (K5) through the C-ABI: bit-exact against the oracle and the cv2 goldens."""
import os

import numpy as np
import pytest

import oracle
from gslam_b200 import synth
from gslam_b200.api import Features

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming_golden.npz"))


def check(ctx, q, t):
    got = ctx.match_hamming(q, t)
    want = oracle.match_hamming(q, t)
    for g, w, name in zip(got, want, ("idx", "d1", "d2")):
        assert np.array_equal(g, w), name


def test_golden_cv2(ctx):
    idx, d1, d2 = ctx.match_hamming(G["q"], G["t"])
    assert np.array_equal(idx, G["idx"]) and np.array_equal(d1, G["d1"]) and np.array_equal(d2, G["d2"])


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (7, 33), (128, 128), (129, 1000), (1000, 1000), (2000, 2000), (513, 4097), (3000, 50)])
def test_vs_oracle(ctx, nq, nt):
    check(ctx, synth.random_descriptors(nq, 100 + nq), synth.random_descriptors(nt, 200 + nt))


def test_ties_and_duplicates(ctx):
    q = synth.random_descriptors(500, 1); t = synth.random_descriptors(900, 2)
    t[::7] = t[3]           # many duplicates spread over all train chunks
    q[::5] = t[3]
    q[1] = t[899]; t[450] = t[899]
    t[600:620] = 0; q[2] = 0
    check(ctx, q, t)
    idx, d1, d2 = ctx.match_hamming(q, t)
    assert idx[0] == 0 and d1[0] == 0 and d2[0] == 0  # t[0]=t[3] duplicate: lowest index wins


def test_empty_and_ragged(ctx):
    q = synth.random_descriptors(10, 3)
    idx, d1, d2 = ctx.match_hamming(q, np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (d1 == 257).all() and (d2 == 257).all()
    idx, d1, d2 = ctx.match_hamming(q, q[:1])
    assert (idx == 0).all() and (d2 == 257).all()
    idx, d1, d2 = ctx.match_hamming(np.zeros((0, 32), np.uint8), q)
    assert idx.size == 0


def test_all_zero_all_one(ctx):
    q = np.zeros((4, 32), np.uint8); t = np.full((3, 32), 255, np.uint8)
    idx, d1, d2 = ctx.match_hamming(q, t)
    assert (d1 == 256).all() and (idx == 0).all() and (d2 == 256).all()


def test_device_resident_features(ctx):
    q = synth.random_descriptors(2000, 9); t = synth.random_descriptors(2000, 10)
    fq, ft = Features(ctx, 2048), Features(ctx, 2048)
    fq.upload(q); ft.upload(t)
    for _ in range(3):  # repeated launches reuse the tickets
        fq.match(ft)
    got = fq.matches()
    want = oracle.match_hamming(q, t)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # symmetric property at full size: matching a set against itself returns the identity with distance 0
    fq.match(fq)
    idx, d1, _ = fq.matches()
    assert np.array_equal(idx, np.arange(2000)) and (d1 == 0).all()
    fq.close(); ft.close()


def _stereo_case(rng, nl, nr, dup=False):
    kl = np.zeros(nl, oracle.KP_DTYPE); kr = np.zeros(nr, oracle.KP_DTYPE)
    kl["x"] = rng.uniform(0, 752, nl).astype(np.float32); kl["y"] = np.round(rng.uniform(0, 480, nl) * 2) / 2
    kr["x"] = rng.uniform(0, 752, nr).astype(np.float32); kr["y"] = np.round(rng.uniform(0, 480, nr) * 2) / 2
    dl = rng.integers(0, 256, (nl, 32), dtype=np.uint8); dr = rng.integers(0, 256, (nr, 32), dtype=np.uint8)
    if dup and nr > 8:
        dr[5] = dr[2]; kr["y"][5] = kr["y"][2]; dr[7] = dr[2]; kr["y"][7] = kr["y"][2]
    return kl, dl, kr, dr


@pytest.mark.gpu
@pytest.mark.parametrize("nl,nr,band,mind,maxd,dup", [(300, 280, 2.0, 0.0, 96.0, False), (64, 500, 0.0, -5.0, 1e9, True), (200, 0, 2.0, 0.0, 50.0, False),
                                                       (1, 1, 1000.0, -1e9, 1e9, False), (2000, 2000, 2.0, 0.0, 120.0, True), (2047, 1025, 2.5, 3.0, 200.0, True)])
def test_stereo_rowband_match_matches_oracle(ctx, nl, nr, band, mind, maxd, dup):
    """gb_match_stereo (row band + disparity window, BASELINE config 4's stereo association) bit-exact vs oracle/hamming_ref.c."""
    rng = np.random.default_rng(nl * 7 + nr)
    kl, dl, kr, dr = _stereo_case(rng, nl, nr, dup)
    got = ctx.match_stereo(kl, dl, kr, dr, band, mind, maxd)
    want = oracle.match_stereo(kl, dl, kr, dr, band, mind, maxd)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.gpu
def test_stereo_match_on_extracted_pair_device_resident(ctx):
    """Left / right frames extracted on the device and associated without a host round trip (extract -> extract -> stereo match
    on one stream, device-side keypoint counts) == the oracle on the downloaded features."""
    from gslam_b200 import synth
    from gslam_b200.api import Features
    left = synth.synth_frame(752, 480, seed=21)
    right = np.roll(left, -14, axis=1)  # a fronto-parallel scene at constant disparity 14 px
    cfg = ctx.orb_cfg(nfeatures=2000)
    fl, fr = Features(ctx, 4256), Features(ctx, 4256)
    fl.extract(left, 752, 480, cfg); fr.extract(right, 752, 480, cfg)
    fl.match_stereo(fr, 2.0, 0.0, 96.0)
    idx, d1, d2 = fl.matches()
    kl, dl = fl.download(); kr, dr = fr.download()
    want = oracle.match_stereo(kl, dl, kr, dr, 2.0, 0.0, 96.0)
    assert np.array_equal(idx, want[0]) and np.array_equal(d1, want[1]) and np.array_equal(d2, want[2])
    ok = idx >= 0
    disp = kl["x"][ok] - kr["x"][idx[ok]]
    assert ok.sum() > 1000 and np.median(np.abs(disp - 14.0)) < 1.0     # the association recovers the disparity
    fl.close(); fr.close()
