import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle
from gslam_b200 import synth
from gslam_b200.api import Context, BAGraph, OptimzeConfig
ctx = Context(0)
a = synth.synth_ba(n_cams=20, n_points=400, obs_per_point=4, n_fixed=2, seed=3, pose_sigma_t=1.0, pose_sigma_deg=10, point_sigma=2.0)
for it in (1, 2, 3, 4, 5, 6, 8, 10, 15):
    b = a.copy()
    r0 = oracle.ba_solve(b, max_iterations=it, function_tolerance=0.0, pcg_max_iters=50, pcg_tol=1e-10)
    out = [f"it={it} oracle acc={r0.accepted} cost={r0.final_cost:.12e} lam={r0.lambda_final:.3e}"]
    for mode in (0, 2, 1):
        c = a.copy(); g = BAGraph(ctx, c); g.force_generic_pcg(mode)
        cf = OptimzeConfig(); cf.maxIterations = it; cf.functionTolerance = 0.0; cf.pcgMaxIterations = 50; cf.pcgTolerance = 1e-10
        r1 = g.solve(cf); g.close()
        out.append(f"m{mode} acc={r1.accepted} cost={r1.final_cost:.12e} lam={r1.lambda_final:.3e}")
    print(" | ".join(out))
