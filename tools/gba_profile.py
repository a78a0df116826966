"""Phase clocks of the block-CSR PCG kernel (ba_pcg_bcsr.cu) and a launch breakdown of one global-BA LM iteration (config 5)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gslam_b200 import capi, synth
from gslam_b200.api import BAGraph, Context, OptimzeConfig

ctx = Context(0)
L = capi.lib()
L.gb_dbg_ba_pcg_profile.restype = C.c_int
L.gb_dbg_ba_pcg_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
pb = synth.synth_ba(500, 100000, 10, seed=42, n_fixed=2)
for env in ({}, {"GB_BA_NO_PCG_CLUSTER": "1"}):
    for k in ("GB_BA_NO_PCG_CLUSTER",):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = BAGraph(ctx, pb)
    cfg = OptimzeConfig(maxIterations=5, functionTolerance=0.0, pcgMaxIterations=30)
    g.solve(cfg)
    out = np.zeros(8, np.int64)
    L.gb_dbg_ba_pcg_profile(ctx._h, g._h, out.ctypes.data_as(C.c_void_p))   # enable
    g.reset(); r = g.solve(cfg)
    L.gb_dbg_ba_pcg_profile(ctx._h, g._h, out.ctypes.data_as(C.c_void_p))   # read (last launch)
    it = max(int(out[6]), 1)
    print("mode", env or "cluster", "gpu_ms", r.gpu_ms, "iters", it, "clk/iter: matvec+dots", out[1] // it, "barrier1", out[2] // it, "update+publish", out[3] // it,
          "barrier2", out[4] // it, "total clk", out[5],
          "| inside the first: own mat-vec", out[0] // it, "wait for the CTA", out[7] // it)
    g.close()
