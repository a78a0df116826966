import ctypes as C, numpy as np, sys
sys.path.insert(0,'.')
from gslam_b200 import capi, synth
from gslam_b200.api import Context, BAGraph, OptimzeConfig
ctx=Context(0); L=capi.lib(); L.gb_dbg_ba_pcg_profile.restype=C.c_int
pb=synth.synth_ba(50,2000,5,seed=42,n_fixed=2); g=BAGraph(ctx,pb)
print('cluster', g.pcg_cluster_size(), 'sparse blocks', g.pcg_sparse_blocks())
out=(C.c_longlong*8)()
L.gb_dbg_ba_pcg_profile(ctx._h, g._h, out)
c=OptimzeConfig(maxIterations=3,functionTolerance=0.0)
g.solve(c); g.reset(); r=g.solve(c)
L.gb_dbg_ba_pcg_profile(ctx._h, g._h, out)
v=list(out); print('iteration-3 stamps', [v[i]-v[0] for i in range(6)], 'setup cycles', v[7], 'loop end - it3 start', v[6]-v[0], 'gpu_ms', r.gpu_ms)
