"""Exactly the bench's device-resident step, a few times, for ncu (launch list / --set full captures).
usage: python tools/profile_step.py [steps] [--global-ba]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gslam_b200 import synth
from gslam_b200.api import BAGraph, Context, Features, OptimzeConfig
W, H, NKP = 1920, 1080, 2000
steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
ctx = Context(0)
if "--global-ba" in sys.argv:
    pb = synth.synth_ba(500, 100000, 10, seed=42, n_fixed=2)
    g = BAGraph(ctx, pb)
    cfg = OptimzeConfig(maxIterations=2, functionTolerance=0.0, pcgMaxIterations=5)
    for _ in range(steps):
        g.reset(); g.solve(cfg)
    ctx.sync(); sys.exit(0)
cfg = ctx.orb_cfg(nfeatures=NKP)
ba_cfg = OptimzeConfig(maxIterations=10, functionTolerance=0.0, pcgMaxIterations=50, pcgTolerance=1e-10)
base = synth.synth_stream(W, H, 4, seed=7)
ring = torch.from_numpy(base).cuda()
feats = [Features(ctx, 2 * NKP + 256), Features(ctx, 2 * NKP + 256)]
graph = BAGraph(ctx, synth.synth_ba(50, 2000, 5, seed=42, n_fixed=2))
feats[1].extract(ring[3].data_ptr(), W, H, cfg, device_ptr=True, pitch=W)
for k in range(steps):
    f, fp = feats[k & 1], feats[(k + 1) & 1]
    f.extract(ring[k % 4].data_ptr(), W, H, cfg, device_ptr=True, pitch=W)
    f.match(fp)
    graph.reset(); graph.solve(ba_cfg)
ctx.sync()
