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
