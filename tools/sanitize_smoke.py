"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck): one 640x480 extraction + match,
one local-BA solve on each PCG path, one PnP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gslam_b200 import synth
from gslam_b200.api import Context, BAGraph, OptimzeConfig

ctx = Context(0)
a = synth.synth_frame(640, 480, seed=1); b = synth.synth_frame(640, 480, seed=2)
k1, d1 = ctx.orb_extract(a, 500); k2, d2 = ctx.orb_extract(b, 500)
idx, dist, dist2 = ctx.match_hamming(d1, d2)
print("orb", len(k1), len(k2), "match", int(dist.min()), int(dist.max()))
cfg = OptimzeConfig(maxIterations=3, functionTolerance=0.0, pcgMaxIterations=20, pcgTolerance=1e-10)
for shape, mode in (((20, 300, 4), 0), ((20, 300, 4), 2), ((20, 300, 4), 1), ((60, 600, 4), 0)):
    pb = synth.synth_ba(*shape, seed=3, n_fixed=2)
    g = BAGraph(ctx, pb); g.force_generic_pcg(mode)
    r = g.solve(cfg); print("ba", shape, mode, r.iterations, r.accepted, f"{r.final_cost:.6e}")
    g.close()
pb = synth.synth_ba(20, 300, 4, seed=3, n_fixed=2)
r = ctx.ba_solve(pb, cfg); print("one-shot", f"{r.final_cost:.6e}")
# ---- round 2 kernels: the large-graph sweep (teams, cp.async stages, bulk-copied W tiles) with the chunked Schur complement and the
# block-CSR PCG around it, pose-graph terms, the BoW transform (walk + single-CTA sort), stereo match, PnP-RANSAC
from gslam_b200.api import Vocabulary
pb = synth.synth_ba(60, 3000, 6, seed=5, n_fixed=2)
g = BAGraph(ctx, pb); g.set_sweep(2)
lin = g.dbg_linearize(0.01); print("sweep2", f"{lin['cost']:.6e}")
g.force_generic_pcg(1); r = g.solve(cfg); print("sweep2 solve", f"{r.final_cost:.6e}")
g.close()
pb = synth.synth_ba(300, 40, all_visible=True, n_fixed=2, seed=5)      # multi-chunk landmarks
g = BAGraph(ctx, pb); g.set_sweep(2); print("sweep2 long landmarks", f"{g.dbg_linearize(0.01)['cost']:.6e}"); g.close()
pb = synth.synth_ba(130, 6000, 8, seed=6, n_fixed=2)                    # chunked Schur + block-CSR PCG
r = ctx.ba_solve(pb, OptimzeConfig(maxIterations=2, functionTolerance=0.0, pcgMaxIterations=20)); print("large path", f"{r.final_cost:.6e}")
pb = synth.synth_ba(20, 300, 4, seed=3, n_fixed=2); pe = synth.synth_pose_edges(pb, seed=1, n_loops=5, gps_every=4, with_info=True)
r = ctx.ba_solve_posegraph(pb, pe, cfg); print("pose graph terms", f"{r.final_cost:.6e}")
vt = synth.synth_vocabulary(10, 3, seed=1, prune=0.1, stop=0.05)
dv = Vocabulary(ctx, vt.k, vt.L, vt.weighting, vt.scoring, vt.child_num, vt.weight, vt.desc)
out = dv.transform(d1, 1); print("bow", len(out["words"]), len(out["fv_feat"]))
big = np.repeat(d1, 20, axis=0)[:9000]; out = dv.transform(big, 0); print("bow (global-memory keys)", len(out["words"]))
dv.close()
