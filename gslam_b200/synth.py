"""Deterministic synthetic inputs for tests and bench.py.

Frames: integer-only generator (no libm, no numpy RNG stream dependence) so the same bytes come out on every box:
mid-gray 64 + random filled rectangles / discs + integer pseudo-Gaussian noise; the next frame of a stream is the
same scene shifted by (3,5) px.  BA graphs: KITTI-00-like planar trajectory (numpy default_rng; fixtures commit the
arrays themselves where bit-identity matters).
"""
from __future__ import annotations

import dataclasses

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser: uint64 -> uint64 (pure integer ops, wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _rand_u64(seed: int, stream: int, n: int) -> np.ndarray:
    base = np.uint64((seed * 0x100000001B3 + stream * 0x9E3779B1 + 0x1234567) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        return _splitmix64(base + np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95))


def synth_scene(width: int, height: int, seed: int = 0, margin: int = 0) -> np.ndarray:
    """Noise-free scene (uint8, (height+2*margin) x (width+2*margin)): gray 64 + ~3000*(W*H/921600) shapes."""
    W, H = width + 2 * margin, height + 2 * margin
    img = np.full((H, W), 64, dtype=np.uint8)
    n_shapes = max(8, int(3000 * (W * H) / 921600))
    r = _rand_u64(seed, 1, n_shapes * 6).reshape(n_shapes, 6)
    cx = (r[:, 0] % np.uint64(W)).astype(np.int64)
    cy = (r[:, 1] % np.uint64(H)).astype(np.int64)
    sw = (r[:, 2] % np.uint64(37)).astype(np.int64) + 4
    sh = (r[:, 3] % np.uint64(37)).astype(np.int64) + 4
    gray = (r[:, 4] % np.uint64(256)).astype(np.uint8)
    kind = (r[:, 5] % np.uint64(2)).astype(np.int64)
    for i in range(n_shapes):
        x0, x1 = max(0, cx[i] - sw[i] // 2), min(W, cx[i] + sw[i] // 2 + 1)
        y0, y1 = max(0, cy[i] - sh[i] // 2), min(H, cy[i] + sh[i] // 2 + 1)
        if x1 <= x0 or y1 <= y0:
            continue
        if kind[i] == 0:
            img[y0:y1, x0:x1] = gray[i]
        else:
            rad = int(sw[i]) // 2
            yy, xx = np.ogrid[y0:y1, x0:x1]
            mask = (xx - cx[i]) ** 2 + (yy - cy[i]) ** 2 <= rad * rad
            img[y0:y1, x0:x1][mask] = gray[i]
    return img


def _noise(width: int, height: int, seed: int) -> np.ndarray:
    """Integer pseudo-Gaussian noise, sigma ~3: sum of 12 uniform{0..5} minus 30 (variance 12*35/12=35 -> sigma 5.9)/2."""
    r = _rand_u64(seed, 2, width * height * 2).reshape(2, height, width)
    acc = np.zeros((height, width), dtype=np.int64)
    for half in range(2):
        v = r[half]
        for k in range(6):
            acc += ((v >> np.uint64(8 * k)) % np.uint64(6)).astype(np.int64)
    return (acc - 30) // 2  # sigma ~ 2.96


def synth_frame(width: int, height: int, seed: int = 0) -> np.ndarray:
    """A single frame (frame 0 of stream `seed`).  uint8 H x W."""
    return synth_stream(width, height, 1, seed)[0]


def synth_stream(width: int, height: int, n_frames: int, seed: int = 0) -> np.ndarray:
    """n_frames consecutive frames (n, H, W) uint8; frame k+1 is frame k's scene moved by (3,5) px."""
    big = synth_scene(width + 3 * n_frames, height + 5 * n_frames, seed)
    out = np.empty((n_frames, height, width), dtype=np.uint8)
    for k in range(n_frames):
        # camera moves so that scene content shifts by (-3,-5): crop window advances by (3,5)
        crop = big[5 * k:5 * k + height, 3 * k:3 * k + width].astype(np.int64)
        out[k] = np.clip(crop + _noise(width, height, seed * 1000003 + k), 0, 255).astype(np.uint8)
    return out


def random_descriptors(n: int, seed: int = 0) -> np.ndarray:
    """n x 32 uint8 pseudo-random descriptors (integer-only generator)."""
    r = _rand_u64(seed, 3, n * 4)
    return r.view(np.uint8).reshape(n, 32).copy()


# ------------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class BAProblem:
    """SoA bundle-adjustment graph, the Python-side mirror of gb_ba_problem (include/gslam_b200.h)."""
    cam_pose_wc: np.ndarray   # (n_cams, 7) f64  qx qy qz qw tx ty tz  (T_wc, GSLAM SE3 layout)
    cam_dof: np.ndarray       # (n_cams,) u8
    points: np.ndarray        # (n_points, 3) f64
    point_free: np.ndarray    # (n_points,) u8
    obs_cam: np.ndarray       # (n_obs,) i32
    obs_point: np.ndarray     # (n_obs,) i32
    obs_xyz: np.ndarray       # (n_obs, 3) f64 CameraAnchor
    obs_info: np.ndarray | None = None  # (n_obs, 4) f64 or None
    gt_pose_wc: np.ndarray | None = None
    gt_points: np.ndarray | None = None

    @property
    def n_cams(self): return int(self.cam_pose_wc.shape[0])
    @property
    def n_points(self): return int(self.points.shape[0])
    @property
    def n_obs(self): return int(self.obs_cam.shape[0])

    def copy(self) -> "BAProblem":
        return BAProblem(**{f.name: (None if getattr(self, f.name) is None else np.array(getattr(self, f.name), copy=True))
                            for f in dataclasses.fields(self)})


def _quat_from_yaw(yaw: np.ndarray) -> np.ndarray:
    """Rotation about the camera y axis (down) by `yaw` rad -> (n,4) x,y,z,w."""
    q = np.zeros((yaw.shape[0], 4))
    q[:, 1] = np.sin(0.5 * yaw)
    q[:, 3] = np.cos(0.5 * yaw)
    return q


def _quat_to_R(q: np.ndarray) -> np.ndarray:
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def synth_ba(n_cams: int = 50, n_points: int = 2000, obs_per_point: int = 5, seed: int = 42, n_fixed: int = 2,
             pixel_sigma: float = 1.0, focal: float = 718.0, pose_sigma_t: float = 0.05, pose_sigma_deg: float = 0.5,
             point_sigma: float = 0.1, all_visible: bool = False) -> BAProblem:
    """KITTI-00-shaped synthetic BA graph.

    Cameras advance 1 m per keyframe along world +z (camera axes: x right, y down, z forward) with +-5 deg yaw jitter;
    every landmark is seen by `obs_per_point` consecutive keyframes (or by all, `all_visible`); measurements are
    normalised projections + N(0,(pixel_sigma/focal)^2); the initial estimate is the truth perturbed by
    N(0,pose_sigma_t) m / N(0,pose_sigma_deg) deg (cameras) and N(0,point_sigma) m (points); the first `n_fixed`
    keyframes are exact and fixed (UPDATE_KF_NONE) — the gauge anchor of a sliding local-BA window.
    """
    rng = np.random.default_rng(seed)
    d = n_cams if all_visible else min(obs_per_point, n_cams)
    spacing = 0.2 if all_visible else 1.0
    yaw = np.deg2rad(rng.uniform(-5.0, 5.0, n_cams))
    q_wc = _quat_from_yaw(yaw)
    t_wc = np.zeros((n_cams, 3))
    t_wc[:, 2] = spacing * np.arange(n_cams)
    t_wc[:, 0] = 0.05 * rng.standard_normal(n_cams)
    R_wc = _quat_to_R(q_wc)

    first = rng.integers(0, n_cams - d + 1, n_points)
    depth = rng.uniform(5.0 + spacing * d, max(50.0, 25.0 + spacing * d), n_points)  # (50 m unless the track is longer)
    lat = rng.uniform(-0.35, 0.35, n_points) * depth
    ver = rng.uniform(-0.15, 0.15, n_points) * depth
    mid = first + d // 2
    p_cam = np.stack([lat, ver, depth], axis=1)
    pts = np.einsum("nij,nj->ni", R_wc[mid], p_cam) + t_wc[mid]
    # keep depth positive w.r.t. every observing camera: points were placed w.r.t. the middle camera at >= 5+d*spacing

    obs_cam = (first[:, None] + np.arange(d)[None, :]).reshape(-1).astype(np.int32)
    obs_point = np.repeat(np.arange(n_points, dtype=np.int32), d)
    pc = np.einsum("nji,nj->ni", R_wc[obs_cam], pts[obs_point] - t_wc[obs_cam])  # R^T (p - t)
    uv = pc[:, :2] / pc[:, 2:3] + (pixel_sigma / focal) * rng.standard_normal((obs_cam.shape[0], 2))
    obs_xyz = np.concatenate([uv, np.ones((uv.shape[0], 1))], axis=1)

    # perturbed initial estimate
    dq = np.zeros((n_cams, 4))
    ang = np.deg2rad(pose_sigma_deg) * rng.standard_normal((n_cams, 3))
    dq[:, :3] = 0.5 * ang
    dq[:, 3] = 1.0
    dq /= np.linalg.norm(dq, axis=1, keepdims=True)
    q0 = _quat_mul(q_wc, dq)
    t0 = t_wc + pose_sigma_t * rng.standard_normal((n_cams, 3))
    q0[:n_fixed] = q_wc[:n_fixed]
    t0[:n_fixed] = t_wc[:n_fixed]
    p0 = pts + point_sigma * rng.standard_normal(pts.shape)
    dof = np.full(n_cams, 63, dtype=np.uint8)
    dof[:n_fixed] = 0
    # shuffle the edge order: callers hand edges over in arbitrary order (BundleGraph::mappointObserves)
    perm = rng.permutation(obs_cam.shape[0])
    return BAProblem(cam_pose_wc=np.ascontiguousarray(np.concatenate([q0, t0], axis=1)), cam_dof=dof,
                     points=np.ascontiguousarray(p0), point_free=np.ones(n_points, dtype=np.uint8),
                     obs_cam=np.ascontiguousarray(obs_cam[perm]), obs_point=np.ascontiguousarray(obs_point[perm]),
                     obs_xyz=np.ascontiguousarray(obs_xyz[perm]), obs_info=None,
                     gt_pose_wc=np.concatenate([q_wc, t_wc], axis=1), gt_points=pts)


# ---- pose-graph terms (GSLAM::SE3Edge / GPSEdge, Optimizer.h:127-148) -------------------------------------------------------------
@dataclasses.dataclass
class PoseEdges:
    """Python-side mirror of gb_pose_edges (include/gslam_b200.h)."""
    se3_first: np.ndarray            # (n_se3,) i32
    se3_second: np.ndarray           # (n_se3,) i32
    se3_meas: np.ndarray             # (n_se3, 7) f64  SE3_12 = SE3_1^-1 SE3_2
    se3_info: np.ndarray | None      # (n_se3, 36) f64 or None
    gps_frame: np.ndarray            # (n_gps,) i32
    gps_meas: np.ndarray             # (n_gps, 7) f64  prior on T_wc
    gps_info: np.ndarray | None      # (n_gps, 36) f64 or None

    @property
    def n_se3(self): return int(self.se3_first.shape[0])
    @property
    def n_gps(self): return int(self.gps_frame.shape[0])


def _quat_rot(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    return np.einsum("...ij,...j->...i", _quat_to_R(q), v)


def se3_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(…,7) x (…,7): rotation product, translation a.R b.t + a.t (SE3.h:129-131)."""
    return np.concatenate([_quat_mul(a[..., :4], b[..., :4]), _quat_rot(a[..., :4], b[..., 4:]) + a[..., 4:]], axis=-1)


def se3_inv(a: np.ndarray) -> np.ndarray:
    qi = a[..., :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    return np.concatenate([qi, -_quat_rot(qi, a[..., 4:])], axis=-1)


def _small_se3(rng, n, sigma_t, sigma_r):
    w = sigma_r * rng.standard_normal((n, 3))
    q = np.concatenate([0.5 * w, np.ones((n, 1))], axis=1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([q, sigma_t * rng.standard_normal((n, 3))], axis=1)


def synth_pose_edges(pb: BAProblem, seed: int = 0, odometry: bool = True, n_loops: int = 0, gps_every: int = 0, sigma_t: float = 0.02,
                     sigma_r: float = 0.002, with_info: bool = False) -> PoseEdges:
    """Pose-graph terms for the trajectory of `pb`, measured on its ground truth with small noise: odometry edges between consecutive
    keyframes, `n_loops` random loop closures, a GPS prior on every `gps_every`-th keyframe; optional random SPD 6x6 information."""
    rng = np.random.default_rng(seed)
    T = pb.gt_pose_wc
    n = T.shape[0]
    first, second = [], []
    if odometry:
        first += list(range(n - 1)); second += list(range(1, n))
    for _ in range(n_loops):
        a, b = rng.choice(n, 2, replace=False)
        first.append(int(a)); second.append(int(b))
    first = np.array(first, np.int32); second = np.array(second, np.int32)
    meas = se3_mul(se3_mul(se3_inv(T[first]), T[second]), _small_se3(rng, first.shape[0], sigma_t, sigma_r)) if first.shape[0] else np.zeros((0, 7))
    gps = np.arange(0, n, gps_every, dtype=np.int32) if gps_every > 0 else np.zeros(0, np.int32)
    gmeas = se3_mul(T[gps], _small_se3(rng, gps.shape[0], 5 * sigma_t, 5 * sigma_r)) if gps.shape[0] else np.zeros((0, 7))

    def spd(m):
        A = rng.standard_normal((m, 6, 6))
        M = np.einsum("nij,nkj->nik", A, A) + 6.0 * np.eye(6)
        M[:, :3, :3] *= 4.0
        return np.ascontiguousarray(M.reshape(m, 36))
    return PoseEdges(first, second, np.ascontiguousarray(meas), spd(first.shape[0]) if with_info else None, gps, np.ascontiguousarray(gmeas),
                     spd(gps.shape[0]) if with_info else None)


# ---- vocabulary tree (GSLAM::Vocabulary's flat layout, Vocabulary.h:583-601) --------------------------------------------------------
@dataclasses.dataclass
class VocabularyTree:
    k: int
    L: int
    weighting: int             # Vocabulary::WeightingType (Vocabulary.h:88-94)
    scoring: int               # Vocabulary::ScoringType (:97-105)
    child_num: np.ndarray      # (n_nodes,) u32: children of node p are rows p*k+1 .. p*k+child_num[p]
    weight: np.ndarray         # (n_nodes,) f32
    desc: np.ndarray           # (n_nodes, 32) u8

    @property
    def n_nodes(self): return int(self.child_num.shape[0])


def synth_vocabulary(k=10, L=4, seed=1, weighting=0, scoring=0, prune=0.0, stop=0.0) -> VocabularyTree:
    """A complete k-ary tree of depth L with 256-bit node descriptors (siblings share most of their bits, so near-ties are common) and
    idf-like weights in (0.2, 9); `prune`: fraction of the inner nodes below the root turned into leaves or given fewer than k children
    (an unbalanced tree), `stop`: fraction of the nodes with weight 0 (stopped words)."""
    rng = np.random.default_rng(seed)
    n = (k ** (L + 1) - 1) // (k - 1)
    inner = (k ** L - 1) // (k - 1)
    child = np.zeros(n, np.uint32)
    child[:inner] = k
    if prune > 0:
        cut = rng.random(inner) < prune
        cut[0] = False
        child[:inner][cut] = 0
        short = rng.random(inner) < prune
        sel = short & ~cut
        child[:inner][sel] = rng.integers(1, k + 1, int(sel.sum()))
    weight = rng.uniform(0.2, 9.0, n).astype(np.float32)
    if stop > 0:
        weight[rng.random(n) < stop] = 0.0
    noise = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    base = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    parent = (np.arange(n) - 1) // k
    parent[0] = 0
    flip = rng.random((n, 32)) < 0.12
    desc = np.where(flip, noise, base[parent]).astype(np.uint8)
    return VocabularyTree(int(k), int(L), int(weighting), int(scoring), child, weight, np.ascontiguousarray(desc))
