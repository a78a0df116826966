"""FIRST thing to run on a B200 in round 2: gb_pnp_ransac (gslam_b200/csrc/pnp.cu, written after round 1's GPU budget was spent)
against the CPU checker oracle/pnp_ref.c on synthetic 2D-3D sets — same winning hypothesis, same inlier mask, pose within 1e-7.
Run it once under `compute-sanitizer --tool memcheck` too, then turn the cases into tests/test_pnp_gpu.py."""
import os, sys, time
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R_)
import numpy as np
import oracle
from gslam_b200.api import Context
from scipy.spatial.transform import Rotation as R

ctx = Context(0)
rng = np.random.default_rng(1)
bad = 0
for (n, of, sig) in [(50, 0.0, 0.0), (200, 0.3, 1 / 718), (2000, 0.5, 1 / 718), (1000, 0.7, 1 / 718), (30, 0.2, 2 / 718), (4, 0.0, 0.0)]:
    for rep in range(5):
        Rg = R.from_rotvec(rng.normal(0, 0.3, 3)).as_matrix(); tg = rng.uniform(-1, 1, 3)
        Xc = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(3, 20, n)])
        Xw = (Xc - tg) @ Rg
        xy = Xc[:, :2] / Xc[:, 2:3] + rng.normal(0, sig, (n, 2))
        no = int(of * n); idx = rng.permutation(n)[:no]
        xy[idx] = np.column_stack([rng.uniform(-1.3, 1.3, no), rng.uniform(-1, 1, no)])
        want = oracle.pnp_ransac(Xw, xy, threshold=4 / 718, confidence=0.99, max_hypotheses=1024, seed=rep + 1)
        t0 = time.perf_counter()
        got = ctx.pnp_ransac(Xw, xy, threshold=4 / 718, confidence=0.99, max_hypotheses=1024, seed=rep + 1)
        dt = time.perf_counter() - t0
        same = (got[2].best_hypothesis == want[2].best_hypothesis and got[2].best_root == want[2].best_root
                and got[2].hypotheses == want[2].hypotheses and np.array_equal(got[1], want[1]) and np.abs(got[0] - want[0]).max() < 1e-7)
        bad += not same
        print("ok      " if same else "MISMATCH", n, of, rep, "hyp", got[2].hypotheses, want[2].hypotheses, "winner", got[2].best_hypothesis,
              want[2].best_hypothesis, "inliers", got[2].inliers_refined, want[2].inliers_refined, f"pose diff {np.abs(got[0] - want[0]).max():.2e}",
              f"{dt * 1e3:.2f} ms")
print("mismatches:", bad)
