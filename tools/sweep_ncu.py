"""One fused BA sweep on the config-5-shaped graph, for `ncu --set full -k regex:ba_linearize_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gslam_b200 import synth
from gslam_b200.api import Context, BAGraph

ctx = Context(0)
pb = synth.synth_ba(500, 100000, 10, seed=42, n_fixed=2)
g = BAGraph(ctx, pb)
for _ in range(3):
    g.sweep(0.01)
ctx.sync()
g.close()
