"""Pins oracle/pnp_ref.c — the CPU checker of the NEXT hot-path row (Estimator::findPnP, P3P + RANSAC,
GSLAM/core/Estimator.h:158-164).  The reference ships no implementation, test or vector for it ("parity unpinned"); what pins the
oracle: known roots, ground-truth poses, cv2.solveP3P solution sets and cv2.solvePnPRansac on the same synthetic data."""
import numpy as np
import pytest

import oracle


def _random_pose(rng, rot_sigma=None):
    from scipy.spatial.transform import Rotation as R
    Rg = (R.random(random_state=int(rng.integers(0, 1 << 31))) if rot_sigma is None else R.from_rotvec(rng.normal(0, rot_sigma, 3))).as_matrix()
    return Rg, rng.uniform(-1, 1, 3)


def test_quartic_known_roots():
    rng = np.random.default_rng(0)
    for _ in range(300):
        r = np.sort(rng.uniform(-3, 3, 4))
        got = oracle.quartic_roots(np.poly(r)[::-1] * rng.uniform(0.5, 2))
        assert len(got) == 4 and np.abs(np.sort(got) - r).max() < 1e-7
        r2 = rng.uniform(-3, 3, 2); z = complex(rng.uniform(-2, 2), rng.uniform(0.1, 2))
        got = oracle.quartic_roots(np.real(np.poly([r2[0], r2[1], z, z.conjugate()]))[::-1])
        assert len(got) == 2 and np.abs(np.sort(got) - np.sort(r2)).max() < 1e-7
    assert len(oracle.quartic_roots([1.0, 0.0, 0.0, 0.0, 1.0])) == 0                       # x^4 + 1
    assert np.allclose(oracle.quartic_roots([0.0, 0.0, 0.0, 0.0, 1.0]), [0.0], atol=1e-12)  # x^4: one (quadruple) root
    dbl = oracle.quartic_roots(np.poly([1.5, 1.5, -2.0, 0.25])[::-1])                       # a double root is reported (once)
    assert np.isclose(dbl, 1.5, atol=1e-6).sum() >= 1 and np.isclose(dbl, -2.0, atol=1e-9).any() and np.isclose(dbl, 0.25, atol=1e-9).any()


def test_p3p_contains_ground_truth_and_every_cv2_solution():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    counts = np.zeros(5, int)
    for _ in range(400):
        Rg, tg = _random_pose(rng); tg = tg + np.array([0, 0, rng.uniform(4, 10)])
        Xc = np.column_stack([rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3), rng.uniform(2, 9, 3)])  # in front of the camera
        Xw = (Xc - tg) @ Rg
        f = Xc / np.linalg.norm(Xc, axis=1, keepdims=True)
        sols = oracle.p3p(Xw, f)
        counts[len(sols)] += 1
        assert any(np.abs(Rs - Rg).max() < 1e-6 and np.abs(ts - tg).max() < 1e-6 for Rs, ts in sols)
        for Rs, ts in sols:  # every returned solution is a rotation that reproduces the three bearings
            Y = Xw @ Rs.T + ts
            assert abs(np.linalg.det(Rs) - 1) < 1e-9 and np.abs(Y / np.linalg.norm(Y, axis=1, keepdims=True) - f).max() < 1e-7
        uv = (Xc[:, :2] / Xc[:, 2:3]).reshape(3, 1, 2)
        _, rv, tv = cv2.solveP3P(Xw.reshape(3, 1, 3), uv, np.eye(3), None, flags=cv2.SOLVEPNP_P3P)
        for r, t in zip(rv, tv):
            Rc = cv2.Rodrigues(r)[0]; tc = t.ravel(); Y = Xw @ Rc.T + tc
            if (Y[:, 2] > 0).all() and np.abs(Y[:, :2] / Y[:, 2:3] - uv.reshape(3, 2)).max() < 1e-9:  # a genuine cv2 solution
                assert any(np.abs(Rs - Rc).max() < 1e-6 and np.abs(ts - tc).max() < 1e-6 for Rs, ts in sols)
    assert counts[0] == 0 and counts[2] > counts[1] > counts[4] > 0  # 1-4 solutions all occur


def test_p3p_degenerate_inputs():
    f = np.eye(3) * 0 + np.array([[0, 0, 1.0]] * 3)
    assert oracle.p3p(np.array([[0, 0, 5.0], [0, 0, 5.0], [1, 0, 5.0]]), f) == []          # repeated point
    assert oracle.p3p(np.array([[0, 0, 5.0], [1, 0, 5.0], [2, 0, 5.0]]), f) == []          # collinear points


def test_sampling_is_counter_based_and_distinct():
    for n in (4, 5, 17, 2000):
        seen = set()
        for h in range(200):
            idx = oracle.pnp_sample(7, h, n)
            assert len(set(idx)) == 3 and all(0 <= i < n for i in idx)
            assert idx == oracle.pnp_sample(7, h, n)  # a pure function of (seed, h, n)
            seen.add(tuple(idx))
        assert len(seen) > (3 if n == 4 else 50)
    assert oracle.pnp_sample(7, 3, 100) != oracle.pnp_sample(8, 3, 100)


def _scene(rng, n, outlier_fraction, sigma):
    Rg, tg = _random_pose(rng, 0.3)
    Xc = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(3, 20, n)])
    Xw = (Xc - tg) @ Rg
    xy = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, sigma, (n, 2))
    bad = rng.permutation(n)[:int(outlier_fraction * n)]
    xy[bad] = np.column_stack([rng.uniform(-1.3, 1.3, bad.size), rng.uniform(-1, 1, bad.size)])
    good = np.ones(n, bool); good[bad] = False
    return Xw, xy, Rg, tg, good


def _angle_deg(Ra, Rb):
    return float(np.degrees(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))))


@pytest.mark.parametrize("n,outliers,sigma,tol_deg,tol_t", [(50, 0.0, 0.0, 1e-5, 1e-6), (200, 0.3, 1 / 718, 0.15, 0.02),
                                                          (2000, 0.5, 1 / 718, 0.1, 0.01), (1000, 0.7, 1 / 718, 0.15, 0.02)])
def test_ransac_recovers_ground_truth_like_cv2(n, outliers, sigma, tol_deg, tol_t):
    cv2 = pytest.importorskip("cv2")
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(n)
    for rep in range(5):
        Xw, xy, Rg, tg, good = _scene(rng, n, outliers, sigma)
        pose, mask, st = oracle.pnp_ransac(Xw, xy, threshold=4 / 718, confidence=0.99, max_hypotheses=2048, seed=rep + 1)
        Rq = R.from_quat(pose[:4]).as_matrix()
        assert _angle_deg(Rq, Rg) < tol_deg and np.linalg.norm(pose[4:] - tg) < tol_t
        assert (mask.astype(bool) == good).mean() > 0.97 and st.inliers_refined >= st.inliers_minimal >= 4
        assert st.hypotheses % 64 == 0 or st.hypotheses == 2048
        ok, rv, tv, _ = cv2.solvePnPRansac(Xw.reshape(-1, 1, 3), xy.reshape(-1, 1, 2), np.eye(3), None, iterationsCount=2048,
                                           reprojectionError=4 / 718, confidence=0.99, flags=cv2.SOLVEPNP_P3P)
        assert ok and _angle_deg(cv2.Rodrigues(rv)[0], Rq) < 2 * tol_deg + 1e-6  # the two estimators agree at the noise level
        again = oracle.pnp_ransac(Xw, xy, threshold=4 / 718, confidence=0.99, max_hypotheses=2048, seed=rep + 1)
        assert np.array_equal(again[0], pose) and np.array_equal(again[1], mask)  # deterministic


def test_ransac_rejects_garbage():
    rng = np.random.default_rng(5)
    Xw = rng.uniform(-5, 5, (100, 3)) + np.array([0, 0, 10.0]); xy = rng.uniform(-1, 1, (100, 2))  # no consistent pose
    with pytest.raises(RuntimeError):
        oracle.pnp_ransac(Xw, xy, threshold=1e-4, max_hypotheses=256)
    with pytest.raises(RuntimeError):
        oracle.pnp_ransac(Xw[:3], xy[:3])  # fewer than four correspondences


# ---- the product's port (gslam_b200/csrc/pnp.cu), instantiated for the HOST: same source as the kernel, no device needed --------
def _product_lib():
    import ctypes as C
    from gslam_b200 import capi
    L = C.CDLL(capi.LIB_PATH)
    L.gb_dbg_pnp_p3p_host.restype = C.c_int
    L.gb_dbg_pnp_p3p_host.argtypes = [C.c_void_p] * 3
    L.gb_dbg_pnp_minimal_host.restype = C.c_int
    L.gb_dbg_pnp_minimal_host.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64, C.c_void_p,
                                          C.POINTER(capi.PnpStats)]
    return L, capi


def test_product_host_instantiation_of_p3p_equals_oracle():
    L, _ = _product_lib()
    rng = np.random.default_rng(3)
    for _ in range(500):
        Rg, tg = _random_pose(rng); tg = tg + np.array([0, 0, rng.uniform(4, 10)])
        Xc = np.column_stack([rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3), rng.uniform(2, 9, 3)])
        Xw = np.ascontiguousarray((Xc - tg) @ Rg); f = np.ascontiguousarray(Xc / np.linalg.norm(Xc, axis=1, keepdims=True))
        out = np.zeros(48)
        ns = L.gb_dbg_pnp_p3p_host(Xw.ctypes.data, f.ctypes.data, out.ctypes.data)
        want = oracle.p3p(Xw, f)
        assert ns == len(want)
        for k, (Rs, ts) in enumerate(want):  # same source, same compiler family, no contraction on either side: bit-identical
            assert np.array_equal(out[12 * k:12 * k + 9].reshape(3, 3), Rs) and np.array_equal(out[12 * k + 9:12 * k + 12], ts)


@pytest.mark.parametrize("n,outliers,sigma", [(50, 0.0, 0.0), (300, 0.4, 1 / 718), (1500, 0.6, 1 / 718), (4, 0.0, 0.0)])
def test_product_host_instantiation_of_the_minimal_stage_equals_oracle(n, outliers, sigma):
    """Sampling, P3P, scoring and the stopping-rule replay of gb_pnp_ransac (host instantiation) pick the oracle's winner."""
    import ctypes as C
    L, capi = _product_lib()
    rng = np.random.default_rng(n + 1)
    for rep in range(4):
        Xw, xy, Rg, tg, good = _scene(rng, n, outliers, sigma)
        Xw = np.ascontiguousarray(Xw); xy = np.ascontiguousarray(xy)
        _, _, want = oracle.pnp_ransac(Xw, xy, threshold=4 / 718, confidence=0.99, max_hypotheses=512, seed=rep + 1)
        Rt = np.zeros(12); st = capi.PnpStats()
        assert L.gb_dbg_pnp_minimal_host(n, Xw.ctypes.data, xy.ctypes.data, 4 / 718, 0.99, 512, rep + 1, Rt.ctypes.data, C.byref(st)) == 0
        assert (st.hypotheses, st.best_hypothesis, st.best_root, st.inliers_minimal) == \
               (want.hypotheses, want.best_hypothesis, want.best_root, want.inliers_minimal)
