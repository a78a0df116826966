"""Pose-graph terms of the BundleGraph on the device (csrc/ba_pose.cu behind gb_ba_graph_create_ex / gb_ba_solve_posegraph:
GSLAM::SE3Edge / GPSEdge, Optimizer.h:127-148) against the oracle (oracle/ba_ref.c, pinned in tests/test_oracle_posegraph.py)."""
import numpy as np
import pytest

import oracle
from oracle import oracle as O
from gslam_b200 import synth
from gslam_b200.api import BAGraph, OptimzeConfig

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def cfg(**kw):
    c = OptimzeConfig()
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def pose_close(a, b, tol):
    s = np.sign(np.sum(a[:, :4] * b[:, :4], axis=1))[:, None]
    assert np.abs(a[:, :4] * s - b[:, :4]).max() < tol
    assert np.abs(a[:, 4:] - b[:, 4:]).max() / max(1.0, np.abs(b[:, 4:]).max()) < tol


CASES = {
    "mixed_info": (dict(n_cams=20, n_points=300, obs_per_point=4, n_fixed=2, seed=3), dict(seed=1, n_loops=5, gps_every=4, with_info=True)),
    "mixed_identity": (dict(n_cams=12, n_points=150, obs_per_point=4, n_fixed=1, seed=8), dict(seed=2, n_loops=3, gps_every=0, with_info=False)),
    "pose_graph_only": (dict(n_cams=40, n_points=0, n_fixed=1, seed=7, pose_sigma_t=0.05, pose_sigma_deg=0.5), dict(seed=2, n_loops=12, gps_every=7, with_info=True)),
    "repeated_pairs": (dict(n_cams=6, n_points=0, n_fixed=1, seed=9), dict(seed=3, n_loops=40, gps_every=2, with_info=True)),  # many edges per pair, both directions
    "gps_only": (dict(n_cams=9, n_points=60, obs_per_point=3, n_fixed=0, seed=4), dict(seed=5, odometry=False, n_loops=0, gps_every=1, with_info=True)),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("mode", [0, 1])
def test_linearisation_and_reduced_system_match_oracle(ctx, name, mode):
    pbk, pek = CASES[name]
    pb = synth.synth_ba(**pbk); pe = synth.synth_pose_edges(pb, **pek)
    want = O.ba_linearize(pb, 0.01, pe)
    g = BAGraph(ctx, pb, pe)
    g.force_generic_pcg(mode)
    got = g.dbg_linearize(0.01); again = g.dbg_linearize(0.01)
    for k in ("U", "gc"):
        assert rel(got[k], want[k]) < 1e-10, k
        assert np.array_equal(got[k], again[k]), k
    assert abs(got["cost"] - want["cost"]) <= 1e-11 * want["cost"]
    S0, gt0, dc0, it0 = O.ba_reduced_system(pb, 0.01, 1e-4, 400, 1e-12, pe)
    S, gt, dc, it = g.dbg_reduced(cfg(pcgMaxIterations=400, pcgTolerance=1e-12))
    assert rel(S, S0) < 1e-9 and rel(gt, gt0) < 1e-9
    assert np.abs(S - S.T).max() < 1e-9 * np.abs(S).max()
    assert rel(dc, dc0) < 1e-6
    g.close()


@pytest.mark.parametrize("name", list(CASES))
def test_solve_matches_oracle(ctx, name):
    pbk, pek = CASES[name]
    pb = synth.synth_ba(**pbk); pe = synth.synth_pose_edges(pb, **pek)
    a = pb.copy(); b = pb.copy()
    # (a pose graph converges to rounding level within a few iterations; beyond that accept / reject is decided by the last bit of
    #  the cost on either side, so the trajectories are compared while the cost still moves)
    iters = 12 if pb.n_points else 4
    kw = dict(max_iterations=iters, function_tolerance=0.0, pcg_max_iters=600, pcg_tol=1e-13)
    r0 = O.ba_solve(a, pe, **kw)
    r1 = ctx.ba_solve_posegraph(b, pe, cfg(maxIterations=iters, functionTolerance=0.0, pcgMaxIterations=600, pcgTolerance=1e-13))
    assert r1.iterations == r0.iterations and r1.accepted == r0.accepted
    assert abs(r1.initial_cost - r0.initial_cost) <= 1e-11 * r0.initial_cost
    assert abs(r1.final_cost - r0.final_cost) <= RTOL * r0.final_cost
    assert r1.final_cost < r1.initial_cost
    pose_close(b.cam_pose_wc, a.cam_pose_wc, RTOL)
    if pb.n_points:
        assert rel(b.points, a.points) < RTOL


def test_rejected_steps_and_default_termination(ctx):
    pb = synth.synth_ba(n_cams=12, n_points=150, obs_per_point=4, n_fixed=2, seed=1, pose_sigma_t=1.0, pose_sigma_deg=10, point_sigma=2.0)
    pe = synth.synth_pose_edges(pb, seed=4, n_loops=4, gps_every=5, with_info=True)
    a = pb.copy(); b = pb.copy()
    r0 = O.ba_solve(a, pe, max_iterations=14, function_tolerance=0.0, pcg_max_iters=600, pcg_tol=1e-13)
    assert r0.accepted < r0.iterations
    r1 = ctx.ba_solve_posegraph(b, pe, cfg(maxIterations=14, functionTolerance=0.0, pcgMaxIterations=600, pcgTolerance=1e-13))
    assert (r1.iterations, r1.accepted) == (r0.iterations, r0.accepted)
    assert abs(r1.final_cost - r0.final_cost) <= RTOL * r0.final_cost
    c = pb.copy()
    r2 = ctx.ba_solve_posegraph(c, pe, cfg())      # reference defaults: stops by the function tolerance
    assert r2.status == 1 and r2.final_cost <= r1.final_cost * 1.01


def test_no_edges_is_the_plain_solve_and_bad_edges_are_refused(ctx):
    pb = synth.synth_ba(10, 200, all_visible=True, n_fixed=2, seed=42)
    a = pb.copy(); b = pb.copy()
    r0 = ctx.ba_solve(a, cfg(maxIterations=5, functionTolerance=0.0))
    empty = synth.synth_pose_edges(pb, odometry=False)
    r1 = ctx.ba_solve_posegraph(b, empty, cfg(maxIterations=5, functionTolerance=0.0))
    assert r0.final_cost == r1.final_cost and np.array_equal(a.cam_pose_wc, b.cam_pose_wc)
    bad = synth.synth_pose_edges(pb, seed=1); bad.se3_second[0] = pb.n_cams
    with pytest.raises(Exception):
        ctx.ba_solve_posegraph(pb.copy(), bad, cfg(maxIterations=1))
    loop = synth.synth_pose_edges(pb, seed=1); loop.se3_second[0] = loop.se3_first[0]
    with pytest.raises(Exception):
        ctx.ba_solve_posegraph(pb.copy(), loop, cfg(maxIterations=1))
    ok = synth.synth_pose_edges(pb, seed=1)
    with pytest.raises(Exception):
        ctx.ba_solve_posegraph(pb.copy(), ok, cfg(maxIterations=1, linearSolver=1))   # the direct solver does not take pose-graph terms
