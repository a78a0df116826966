"""ctypes bindings of oracle/liboracle.so (our C restatements) and oracle/_ref/libgslam_ref.so (the unmodified
reference headers behind an extern "C" shim).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class BaProblemC(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_points", C.c_int32), ("n_obs", C.c_int32),
                ("cam_pose_wc", f64p), ("cam_dof", u8p), ("points", f64p), ("point_free", u8p),
                ("obs_cam", i32p), ("obs_point", i32p), ("obs_xyz", f64p), ("obs_info", f64p)]


class BaOptionsC(C.Structure):
    _fields_ = [("projection", C.c_int32), ("huber_delta", C.c_double), ("max_iterations", C.c_int32),
                ("verbose", C.c_int32), ("function_tolerance", C.c_double), ("lambda_init", C.c_double),
                ("pcg_max_iters", C.c_int32), ("pcg_tol", C.c_double)]


class BaResultC(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("iterations", C.c_int32),
                ("accepted", C.c_int32), ("pcg_iterations", C.c_int32), ("status", C.c_int32),
                ("lambda_final", C.c_double), ("gpu_ms", C.c_float)]


class KeyPointC(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


class OrbCfgC(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("edge_threshold", C.c_int32), ("first_level", C.c_int32), ("wta_k", C.c_int32),
                ("score_type", C.c_int32), ("patch_size", C.c_int32), ("fast_threshold", C.c_int32)]


def default_ba_options(**kw) -> BaOptionsC:
    o = BaOptionsC(0, 0.01, 500, 0, 1e-6, 1e-4, 50, 1e-10)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_orb_cfg(**kw) -> OrbCfgC:
    o = OrbCfgC(500, 1.2, 8, 31, 0, 2, 0, 31, 20)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and _ref/libgslam_ref.so (only where /root/reference exists)."""
    srcs = [os.path.join(_HERE, f) for f in ("hamming_ref.c", "ba_ref.c", "orb_ref.c", "Makefile")]
    lib = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so", "CC=gcc"])
    ref = os.path.join(_HERE, "_ref", "libgslam_ref.so")
    if os.path.isdir("/root/reference/GSLAM/core") and (
            force or not os.path.exists(ref) or os.path.getmtime(os.path.join(_HERE, "ref_shim.cpp")) > os.path.getmtime(ref)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "CXX=g++"])


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_hamming256.restype = C.c_int
        L.orc_hamming256.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_match_hamming.restype = None
        L.orc_match_hamming.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ba_solve.restype = C.c_int
        L.orc_ba_solve.argtypes = [C.POINTER(BaProblemC), C.POINTER(BaOptionsC), C.POINTER(BaResultC)]
        L.orc_ba_linearize.restype = C.c_int
        L.orc_ba_linearize.argtypes = [C.POINTER(BaProblemC), C.c_double] + [C.c_void_p] * 6
        L.orc_ba_cost.restype = C.c_int
        L.orc_ba_cost.argtypes = [C.POINTER(BaProblemC), C.c_double, f64p]
        L.orc_ba_reduced_system.restype = C.c_int
        L.orc_ba_reduced_system.argtypes = [C.POINTER(BaProblemC), C.c_double, C.c_double, C.c_int, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_ba_pnp.restype = C.c_int
        L.orc_ba_pnp.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                 C.POINTER(BaOptionsC), C.POINTER(BaResultC)]
        L.orc_se3_inverse.restype = None
        L.orc_se3_inverse.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_se3_retract.restype = None
        L.orc_se3_retract.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libgslam_ref.so"))


def ref() -> C.CDLL:
    global _REF
    if _REF is None:
        build()
        R = C.CDLL(os.path.join(_HERE, "_ref", "libgslam_ref.so"))
        R.ref_hamming32.restype = C.c_float
        R.ref_hamming32.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("ref_se3_inverse", "ref_se3_exp", "ref_se3_log"):
            getattr(R, name).restype = None
            getattr(R, name).argtypes = [C.c_void_p, C.c_void_p]
        for name in ("ref_se3_transform", "ref_se3_mul"):
            getattr(R, name).restype = None
            getattr(R, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_sim3_raw.restype = None
        R.ref_sim3_raw.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        R.ref_sizeof.restype = C.c_int
        R.ref_sizeof.argtypes = [C.c_char_p]
        R.ref_keypoint_offsets.restype = C.c_int
        R.ref_keypoint_offsets.argtypes = [C.c_void_p]
        _REF = R
    return _REF


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- Hamming -----------------------------------------------------------------------------------------------------
def match_hamming(query: np.ndarray, train: np.ndarray):
    q = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(train, dtype=np.uint8).reshape(-1, 32)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.empty(nq, np.int32); d1 = np.empty(nq, np.int32); d2 = np.empty(nq, np.int32)
    lib().orc_match_hamming(_p(q), nq, _p(t), nt, _p(idx), _p(d1), _p(d2))
    return idx, d1, d2


# ---- BA ------------------------------------------------------------------------------------------------------------
def ba_problem_c(pb) -> tuple[BaProblemC, list]:
    """pb: gslam_b200.synth.BAProblem (arrays are used in place; poses/points are IN/OUT)."""
    keep = []

    def arr(a, dt):
        if a is None:
            return None
        b = np.ascontiguousarray(a, dtype=dt)
        keep.append(b)
        return b
    pose = pb.cam_pose_wc; pts = pb.points
    assert pose.dtype == np.float64 and pose.flags.c_contiguous and pts.dtype == np.float64 and pts.flags.c_contiguous
    dof = arr(pb.cam_dof, np.uint8); pf = arr(pb.point_free, np.uint8)
    oc = arr(pb.obs_cam, np.int32); op = arr(pb.obs_point, np.int32); ox = arr(pb.obs_xyz, np.float64); oi = arr(pb.obs_info, np.float64)
    c = BaProblemC(pb.n_cams, pb.n_points, pb.n_obs, pose.ctypes.data_as(f64p),
                   None if dof is None else dof.ctypes.data_as(u8p), pts.ctypes.data_as(f64p),
                   None if pf is None else pf.ctypes.data_as(u8p), oc.ctypes.data_as(i32p), op.ctypes.data_as(i32p),
                   ox.ctypes.data_as(f64p), None if oi is None else oi.ctypes.data_as(f64p))
    return c, keep


def ba_solve(pb, **opts):
    c, keep = ba_problem_c(pb)
    o = default_ba_options(**opts)
    r = BaResultC()
    rc = lib().orc_ba_solve(C.byref(c), C.byref(o), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"orc_ba_solve failed rc={rc}")
    return r


def ba_linearize(pb, delta=0.01):
    c, keep = ba_problem_c(pb)
    U = np.zeros((pb.n_cams, 6, 6)); gc = np.zeros((pb.n_cams, 6)); V = np.zeros((pb.n_points, 3, 3)); gp = np.zeros((pb.n_points, 3))
    W = np.zeros((pb.n_obs, 6, 3)); cost = np.zeros(1)
    rc = lib().orc_ba_linearize(C.byref(c), delta, _p(U), _p(gc), _p(V), _p(gp), _p(W), _p(cost))
    assert rc == 0
    return dict(U=U, gc=gc, V=V, gp=gp, W=W, cost=float(cost[0]))


def ba_cost(pb, delta=0.01) -> float:
    c, keep = ba_problem_c(pb)
    out = C.c_double()
    rc = lib().orc_ba_cost(C.byref(c), delta, C.byref(out))
    assert rc == 0
    return out.value


def ba_reduced_system(pb, delta=0.01, lam=1e-4, pcg_max_iters=50, pcg_tol=1e-10):
    c, keep = ba_problem_c(pb)
    n6 = 6 * pb.n_cams
    S = np.zeros((n6, n6)); gt = np.zeros(n6); dc = np.zeros(n6); it = C.c_int()
    rc = lib().orc_ba_reduced_system(C.byref(c), delta, lam, pcg_max_iters, pcg_tol, _p(S), _p(gt), _p(dc), C.byref(it))
    assert rc == 0
    return S, gt, dc, it.value


def ba_pnp(xyz, xy1, pose_wc, dof=63, want_info=False, **opts):
    xyz = np.ascontiguousarray(xyz, np.float64); xy1 = np.ascontiguousarray(xy1, np.float64)
    pose = np.ascontiguousarray(pose_wc, np.float64).copy()
    info = np.zeros((6, 6)) if want_info else None
    o = default_ba_options(**opts); r = BaResultC()
    rc = lib().orc_ba_pnp(xyz.shape[0], _p(xyz), _p(xy1), _p(pose), dof, _p(info), C.byref(o), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"orc_ba_pnp failed rc={rc}")
    return pose, r, info
