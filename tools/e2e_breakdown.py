"""Wall-clock breakdown of the end-to-end (host-buffer) step of bench.py: where the time beyond the kernels goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gslam_b200 import synth
from gslam_b200.api import Context, OptimzeConfig

W, H, NKP = 1920, 1080, 2000
ctx = Context(0)
frames = [torch.from_numpy(synth.synth_frame(W, H, seed=k)).pin_memory().numpy() for k in range(4)]
pb0 = synth.synth_ba(50, 2000, 5, seed=42, n_fixed=2)
cfg = OptimzeConfig(maxIterations=10, functionTolerance=0.0, pcgMaxIterations=50, pcgTolerance=1e-10)
prev = ctx.orb_extract(frames[0], NKP)[1]
acc = dict(extract=0.0, match=0.0, copy=0.0, ba=0.0)
N = 40
for k in range(N + 3):
    t0 = time.perf_counter(); kps, desc = ctx.orb_extract(frames[(k + 1) % 4], NKP)
    t1 = time.perf_counter(); ctx.match_hamming(desc, prev); prev = desc
    t2 = time.perf_counter(); pb = pb0.copy()
    t3 = time.perf_counter(); r = ctx.ba_solve(pb, cfg)
    t4 = time.perf_counter()
    if k >= 3:
        acc["extract"] += t1 - t0; acc["match"] += t2 - t1; acc["copy"] += t3 - t2; acc["ba"] += t4 - t3
print({k: round(v / N * 1e3, 4) for k, v in acc.items()}, "ms per call; BA gpu_ms", r.gpu_ms)
