"""Bag-of-words transform on the device (csrc/bow.cu behind gb_voc_create / gb_bow_transform) against the oracle (oracle/bow_ref.c,
itself pinned to the reference's Vocabulary::transform) and the reference-made golden vectors: words, nodes and feature indices
bit-exact; values bit-exact (same float operations in the same order; the double norm is exact in any order for these weights)."""
import os

import numpy as np
import pytest

import oracle
from oracle import oracle as O
from gslam_b200.api import Vocabulary, Features
from gslam_b200 import capi, synth

pytestmark = pytest.mark.gpu
KEYS = ("words", "values", "fv_node", "fv_feat")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bow_golden.npz")


def dev_voc(ctx, v):
    return Vocabulary(ctx, v.k, v.L, v.weighting, v.scoring, v.child_num, v.weight, v.desc)


def queries(v, n, seed, flip=0.05):
    rng = np.random.default_rng(seed)
    src = v.desc[rng.integers(1, v.n_nodes, n)]
    return src ^ np.packbits(rng.random((n, 256)) < flip, axis=1)


def check(got, want):
    for k in KEYS:
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize("q", ["a", "b"])
@pytest.mark.parametrize("lu", [0, 2])
def test_reference_golden_vectors(ctx, q, lu):
    z = np.load(GOLD)
    dv = Vocabulary(ctx, int(z["k"]), int(z["L"]), int(z["weighting"]), int(z["scoring"]), z["child_num"], z["weight"], z["desc"])
    got = dv.transform(z[f"q_{q}"], lu)
    for k in KEYS:
        assert np.array_equal(got[k], z[f"{q}_lu{lu}_{k}"]), k
    dv.close()


@pytest.mark.parametrize("weighting", [O.W_TF_IDF, O.W_TF, O.W_IDF, O.W_BINARY])
@pytest.mark.parametrize("scoring", [O.S_L1, O.S_L2, O.S_KL, O.S_DOT_PRODUCT])
def test_weightings_and_scorings_match_oracle(ctx, weighting, scoring):
    v = O.synth_vocabulary(10, 4, seed=7, weighting=weighting, scoring=scoring, stop=0.1)
    dv = dev_voc(ctx, v)
    f = queries(v, 2000, seed=weighting * 10 + scoring)
    for lu in (0, 1, 4, 6):
        check(dv.transform(f, lu), O.bow_transform(v, f, lu))
    dv.close()


@pytest.mark.parametrize("k,L,n", [(10, 5, 2000), (16, 3, 1000), (20, 3, 777), (32, 2, 4096), (2, 9, 300), (7, 4, 1), (10, 4, 9000)])
def test_tree_shapes_and_sizes(ctx, k, L, n):
    """k <= 16 runs 16 lanes per descriptor, k <= 32 a full warp; 9000 descriptors exceed the shared-memory sort (global-memory keys)."""
    v = O.synth_vocabulary(k, L, seed=k * 100 + L, prune=0.1, stop=0.05)
    dv = dev_voc(ctx, v)
    f = queries(v, n, seed=n)
    for lu in (0, 2):
        check(dv.transform(f, lu), O.bow_transform(v, f, lu))
    dv.close()


def test_ties_take_the_first_child_and_everything_on_one_word(ctx):
    v = O.synth_vocabulary(10, 3, seed=3)
    v.desc[1:11] = v.desc[1]            # ten-way tie at the root for every query
    dv = dev_voc(ctx, v)
    f = queries(v, 500, seed=4)
    got = dv.transform(f, 0); want = O.bow_transform(v, f, 0)
    check(got, want)
    assert np.all((got["words"] >= 111) & (got["words"] < 211))   # all under child 1 of the root (leaves 111..210 are its grandchildren)
    same = np.repeat(f[:1], 1500, axis=0)                        # one word hit 1500 times: the float sum runs over 1500 addends
    check(dv.transform(same, 0), O.bow_transform(v, same, 0))
    dv.close()


def test_empty_input_and_bad_arguments(ctx):
    v = O.synth_vocabulary(10, 2, seed=1)
    dv = dev_voc(ctx, v)
    e = dv.transform(np.zeros((0, 32), np.uint8), 0)
    assert all(e[k].size == 0 for k in KEYS)
    dv.close()
    with pytest.raises(Exception):
        Vocabulary(ctx, 40, 2, 0, 0, v.child_num, v.weight, v.desc)           # k > 32
    bad = v.child_num.copy(); bad[-1] = 3                                      # children beyond the node array
    with pytest.raises(Exception):
        Vocabulary(ctx, v.k, v.L, 0, 0, bad, v.weight, v.desc)


def test_chained_after_extract_without_host_round_trip(ctx):
    """Descriptors still resident in HBM after gb_orb_extract_to (count unknown to the host) go straight into the walk."""
    img = synth.synth_frame(640, 480, seed=3)
    cfg = capi.OrbCfg(); capi.lib().gb_orb_cfg_default(cfg); cfg.nfeatures = 500
    feats = Features(ctx, 1500)
    feats.extract(img, 640, 480, cfg)
    v = O.synth_vocabulary(10, 4, seed=11)
    dv = dev_voc(ctx, v)
    got = dv.transform(feats, 1)
    kps, desc = feats.download()
    assert desc.shape[0] > 100
    check(got, O.bow_transform(v, desc, 1))
    feats.close(); dv.close()
