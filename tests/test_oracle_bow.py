"""Bag-of-words transform (GSLAM/core/Vocabulary.h:1558-1736): the oracle's restatement (oracle/bow_ref.c)
against golden vectors produced by the reference itself (tests/golden/make_golden_bow.py) and, where oracle/_ref is built, against
the live reference -- Vocabulary::create-trained and Vocabulary::load-ed trees, every weighting and scoring type."""
import os

import numpy as np
import pytest

import oracle
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bow_golden.npz")
KEYS = ("words", "values", "fv_node", "fv_feat")


def _need_ref():
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")


def golden_vocabulary():
    z = np.load(GOLD)
    return z, O.VocabularyArrays(int(z["k"]), int(z["L"]), int(z["weighting"]), int(z["scoring"]), z["child_num"], z["weight"], z["desc"])


def queries(v, n, seed, flip=0.05):
    """descriptors a few bit flips away from random nodes of the tree (so that the walk is decided by small distances, ties included)"""
    rng = np.random.default_rng(seed)
    src = v.desc[rng.integers(1, v.n_nodes, n)]
    return src ^ np.packbits(rng.random((n, 256)) < flip, axis=1)


@pytest.mark.parametrize("q", ["a", "b"])
@pytest.mark.parametrize("lu", [0, 2])
def test_oracle_equals_reference_golden(q, lu):
    z, v = golden_vocabulary()
    got = O.bow_transform(v, z[f"q_{q}"], lu)
    for k in KEYS:
        assert np.array_equal(got[k], z[f"{q}_lu{lu}_{k}"]), k
    assert abs(float(got["values"].sum()) - 1.0) < 1e-5    # L1-normalised
    assert np.all(np.diff(got["words"]) > 0)               # std::map order


def test_golden_vocabulary_is_a_trained_tree():
    z, v = golden_vocabulary()
    assert v.k == 8 and v.L == 3 and v.n_nodes == (8 ** 4 - 1) // 7
    leaves = v.child_num == 0
    assert leaves.sum() > 100 and np.all(v.weight[leaves] >= 0)
    # the export is loadable by the reference's own binary loader and walks identically
    if oracle.have_ref():
        R = O.RefVocabulary.from_arrays(v)
        r = R.transform(z["q_a"], 1)
        g = O.bow_transform(v, z["q_a"], 1)
        assert all(np.array_equal(g[k], r[k]) for k in KEYS)
        R.close()


@pytest.mark.parametrize("weighting", [O.W_TF_IDF, O.W_TF, O.W_IDF, O.W_BINARY])
@pytest.mark.parametrize("scoring", [O.S_L1, O.S_L2, O.S_CHI_SQUARE, O.S_KL, O.S_BHATTACHARYYA, O.S_DOT_PRODUCT])
def test_every_weighting_and_scoring_against_live_reference(weighting, scoring):
    _need_ref()
    v = O.synth_vocabulary(10, 3, seed=7, weighting=weighting, scoring=scoring, stop=0.1)
    R = O.RefVocabulary.from_arrays(v)
    f = queries(v, 700, seed=weighting * 10 + scoring)
    for lu in (0, 1, 3, 5):
        want = R.transform(f, lu); got = O.bow_transform(v, f, lu)
        for k in KEYS:
            assert np.array_equal(got[k], want[k]), (k, lu)
    R.close()


def test_trained_vocabulary_against_live_reference():
    _need_ref()
    rng = np.random.default_rng(5)
    centres = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    train = centres[rng.integers(0, 300, (40, 200))] ^ np.packbits(rng.random((40, 200, 256)) < 0.06, axis=2)
    R = O.RefVocabulary.train(train, 40, 10, 3)
    v = R.arrays()
    assert v.n_nodes == 1111
    f = queries(v, 1000, seed=1)
    for lu in (0, 1, 2):
        want = R.transform(f, lu); got = O.bow_transform(v, f, lu)
        for k in KEYS:
            assert np.array_equal(got[k], want[k]), (k, lu)
    for i in range(0, 1000, 97):   # the single-descriptor walk
        w, val, node = R.transform_one(f[i], 1)
        assert (w, node) == (int(got["f_word"][i]), int(O.bow_transform(v, f[i:i + 1], 1)["f_node"][0]))
    R.close()


def test_unbalanced_tree_and_ties():
    """Pruned trees (leaves above level L, inner nodes with fewer than k children) and exact distance ties (first child wins)."""
    _need_ref()
    v = O.synth_vocabulary(10, 4, seed=3, prune=0.15, stop=0.05)
    v.desc[11:21] = v.desc[11]           # the ten children of node 1 are identical: every query reaching node 1 ties ten ways
    R = O.RefVocabulary.from_arrays(v)
    f = queries(v, 1500, seed=9)
    got = O.bow_transform(v, f, 0); want = R.transform(f, 0)
    assert np.array_equal(got["words"], want["words"]) and np.array_equal(got["values"], want["values"])
    under1 = got["f_word"][(got["f_word"] >= 11) & (got["f_word"] <= 20)]
    assert under1.size == 0 or np.all(under1 == 11)
    # levelsup >= L: every feature files under the root (Vocabulary.h:1699-1700)
    top = O.bow_transform(v, f, 4); wtop = R.transform(f, 4)
    assert np.all(top["fv_node"] == 0) and np.array_equal(top["fv_node"], wtop["fv_node"]) and np.array_equal(top["fv_feat"], wtop["fv_feat"])
    # our definition where the reference reads an uninitialised nid: a leaf above the requested level files under itself
    shallow = v.child_num[got["f_word"]] == 0
    assert shallow.all()
    early = got["f_word"] < (10 ** 4 - 1) // 9    # leaves above level 4
    assert early.any() and np.array_equal(got["f_node"][early], got["f_word"][early])
    R.close()


def test_empty_and_single_inputs():
    z, v = golden_vocabulary()
    e = O.bow_transform(v, np.zeros((0, 32), np.uint8), 0)
    assert all(e[k].size == 0 for k in KEYS)
    one = O.bow_transform(v, z["q_a"][:1], 0)
    assert one["words"].size == 1 and one["values"][0] == np.float32(1.0) and one["fv_feat"].tolist() == [0]
