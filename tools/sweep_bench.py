"""Times the BA residual+Jacobian sweep on the config-5-shaped graph (500 cams / 100k landmarks / 1M observations): the latency-tuned
kernel of ba.cu (mode 1) against the bandwidth-tuned persistent kernel of ba_sweep.cu (mode 2), whole and by part."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gslam_b200 import capi, synth
from gslam_b200.api import Context, BAGraph

ctx = Context(0)
L = capi.lib()
for shape in ((500, 100000, 10), (500, 200000, 5), (200, 20000, 10), (50, 2000, 5)):
    pb = synth.synth_ba(*shape, seed=42, n_fixed=2)
    g = BAGraph(ctx, pb)
    b = 168 * pb.n_obs + 96 * pb.n_points + 272 * pb.n_cams
    for mode in (1, 2):
        g.set_sweep(mode)
        g.sweep(0.01); ctx.sync()
        best = 1e9
        for _ in range(5):
            ctx.timer_begin()
            for r in range(20):
                g.sweep(0.01)
            best = min(best, ctx.timer_end() / 20)
        parts = []
        for which in (1, 2):
            L.gb_dbg_ba_sweep_part(ctx._h, g._h, which); ctx.sync()
            ctx.timer_begin()
            for r in range(20):
                L.gb_dbg_ba_sweep_part(ctx._h, g._h, which)
            parts.append(ctx.timer_end() / 20 * 1e3)
        print(f"{shape} mode {mode}: {best * 1e3:.1f} us per sweep, {b / best / 1e6:.0f} GB/s algorithmic; camera pass alone {parts[0]:.1f} us, landmark pass alone {parts[1]:.1f} us", flush=True)
    g.close()
