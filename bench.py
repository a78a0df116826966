#!/usr/bin/env python
"""bench.py — frames/s through detect -> match -> local BA at 1920x1080 mono (BASELINE.json metric), one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one frame through the whole hot path: ORB extract (2000 kp) of a NEW 1080p frame, 256-bit Hamming match against
the previous frame's descriptors, and one local bundle adjustment (50 keyframes / 2000 landmarks / 10 000 observations,
10 LM iterations, 50-iteration block-Jacobi PCG cap).  `value` is measured with every input already resident in HBM
(a ring of frames larger than L2); `e2e` goes through the C-ABI host-buffer entry points (pinned host frame in, keypoints
/ descriptors / matches / poses out, H2D + D2H inside the timed region).  N>1: independent replicas, one rank per GPU
(frames and windows shard with no data-path collective) -> weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NKP = 1920, 1080, 2000
BA_CAMS, BA_PTS, BA_OBS_PER_PT, BA_ITERS, PCG_ITERS = 50, 2000, 5, 10, 50
RING = 72  # frames resident in HBM: 72 * 2.07 MB = 149 MB > 126 MB L2


def level_sizes(w, h, nlevels=8, sf=1.2):
    out = []
    for l in range(nlevels):
        inv = np.float32(1.0) / np.float32(np.float64(np.float32(sf)) ** l)  # cv2: cols * (1/scale) in float, half-even
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes():
    """SURVEY.md §8d per-unit figures."""
    lv = level_sizes(W, H)
    p0 = lv[0][0] * lv[0][1]
    prest = sum(a * b for a, b in lv[1:])
    b_ext = p0 + 2 * prest + NKP * (43 * 43 + 60) + 2 * NKP * 81
    b_match = 32 * (NKP + NKP) + 12 * NKP
    n_obs = BA_PTS * BA_OBS_PER_PT
    b_ba = 168 * n_obs + 96 * BA_PTS + 272 * BA_CAMS
    return dict(extract=b_ext, match=b_match, ba_sweep=b_ba)


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_path(frames, ba_problem, steps, threads):
    """The CPU arm: OpenCV ORB + BFMatcher (the reference's external CPU dependency, when importable; else the oracle
    restatement) + the oracle's ba_ref.  Returns (frames_per_s, description)."""
    import oracle
    try:
        import cv2
        cv2.setNumThreads(threads)
        orb = cv2.ORB_create(nfeatures=NKP)
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)

        def extract(img):
            return orb.detectAndCompute(img, None)[1]

        def match(a, b):
            return bf.match(a, b)
        what = f"cv2 {cv2.__version__} ORB+BFMatcher ({threads} threads)"
    except Exception:
        os.environ["OMP_NUM_THREADS"] = str(threads)

        def extract(img):
            return oracle.orb_extract(img, NKP)[1]

        def match(a, b):
            return oracle.match_hamming(a, b)
        what = f"oracle orb_ref+hamming_ref (OpenMP {threads} threads)"
    prev = extract(frames[0])
    t0 = time.perf_counter()
    for k in range(steps):
        d = extract(frames[(k + 1) % len(frames)])
        match(d, prev)
        prev = d
        pb = ba_problem.copy()
        oracle.ba_solve(pb, max_iterations=BA_ITERS, function_tolerance=0.0, pcg_max_iters=PCG_ITERS, pcg_tol=1e-10)
    dt = time.perf_counter() - t0
    return steps / dt, what + " + oracle ba_ref (1 thread)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    from gslam_b200 import synth
    ba_problem = synth.synth_ba(BA_CAMS, BA_PTS, BA_OBS_PER_PT, seed=42, n_fixed=2)
    cores = os.cpu_count() or 1
    workload = (f"{W}x{H} mono, {NKP} kp ORB extract + {NKP}x{NKP} Hamming match + local BA "
                f"({BA_CAMS} KF/{BA_PTS} pts/{BA_PTS * BA_OBS_PER_PT} obs, {BA_ITERS} LM it, PCG cap {PCG_ITERS})")

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 20))
        frames = synth.synth_stream(W, H, 4, seed=7)
        for _ in range(max(1, min(args.warmup, 2))):
            cpu_path(frames, ba_problem, 1, cores)
        fps, what = cpu_path(frames, ba_problem, steps, cores)
        line = {"impl": "reference", "metric": "frames/sec detect+match+local-BA @1920x1080 mono", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": 1e3 / fps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
                "config": {"workload": workload},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"{steps} frames: {what}"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gslam_b200.api import BAGraph, Context, Features, OptimzeConfig
    ctx = Context(local)
    cfg = ctx.orb_cfg(nfeatures=NKP)
    ba_cfg = OptimzeConfig(maxIterations=BA_ITERS, functionTolerance=0.0, pcgMaxIterations=PCG_ITERS, pcgTolerance=1e-10)

    # ---- device-resident inputs: a ring of distinct frames larger than L2 ----------------------------------------------
    base = synth.synth_stream(W, H, 8, seed=7 + rank)
    ring = torch.empty((RING, H, W), dtype=torch.uint8, device="cuda")
    for k in range(RING):  # distinct content per slot: shifted copies of 8 generated frames (cheap, still cold in L2)
        ring[k] = torch.from_numpy(np.roll(base[k % 8], shift=(k // 8) * 7, axis=1)).cuda()
    feats = [Features(ctx, 2 * NKP + 256), Features(ctx, 2 * NKP + 256)]
    graph = BAGraph(ctx, ba_problem)
    torch.cuda.synchronize()

    def step_device(k):
        f, fp = feats[k & 1], feats[(k + 1) & 1]
        f.extract(ring[k % RING].data_ptr(), W, H, cfg, device_ptr=True, pitch=W)
        f.match(fp)
        graph.reset()
        graph.solve(ba_cfg)

    feats[1].extract(ring[RING - 1].data_ptr(), W, H, cfg, device_ptr=True, pitch=W)
    for k in range(args.warmup):
        step_device(k)
    ctx.sync()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    ctx.timer_begin()
    for k in range(args.steps):
        step_device(args.warmup + k)
    ms = ctx.timer_end()
    torch.cuda.synchronize()
    launches = ctx.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        dist.barrier()
    clocks = sampler.finish() if sampler else None

    # ---- per-stage device timing (CUDA events on the ctx stream) ------------------------------------------------------------
    def time_stage(fn, reps):
        fn(); ctx.sync()
        ctx.timer_begin()
        for r in range(reps):
            fn(r)
        return ctx.timer_end() / reps
    reps = 50
    t_ext = time_stage(lambda r=0: feats[0].extract(ring[(r * 7 + 3) % RING].data_ptr(), W, H, cfg, device_ptr=True, pitch=W), reps)
    t_match = time_stage(lambda r=0: feats[0].match(feats[1]), reps)
    def ba_once(r=0):
        graph.reset(); graph.solve(ba_cfg)
    t_ba = time_stage(ba_once, 10)
    t_sweep_local = time_stage(lambda r=0: graph.sweep(0.01), 50)
    # the BASELINE metric's second clause, "BA Jacobian-eval HBM GB/s": the fused residual+Jacobian sweep (K6a+K6b) on the
    # config-5-shaped graph (500 cams / 100k landmarks / 1M observations: 177.7 MB algorithmic per sweep > L2, so every
    # repetition is cold).  Rank 0 only.
    t_sweep_big, big_bytes = None, None
    if rank == 0:
        big = synth.synth_ba(500, 100000, 10, seed=42, n_fixed=2)
        gbig = BAGraph(ctx, big)
        t_sweep_big = time_stage(lambda r=0: gbig.sweep(0.01), 20)
        big_bytes = 168 * big.n_obs + 96 * big.n_points + 272 * big.n_cams
        gbig.close()
        del big

    # ---- end to end through the host-buffer C-ABI ------------------------------------------------------------------------
    host_frames = [torch.from_numpy(base[k]).pin_memory() for k in range(8)]
    hf = [x.numpy() for x in host_frames]
    e2e_steps = max(10, args.steps // 4)
    prev_desc = ctx.orb_extract(hf[0], NKP)[1]
    def step_host(k):
        nonlocal prev_desc
        kps, desc = ctx.orb_extract(hf[(k + 1) % 8], NKP)
        ctx.match_hamming(desc, prev_desc)
        prev_desc = desc
        pb = ba_problem.copy()
        ctx.ba_solve(pb, ba_cfg)
    for k in range(3):
        step_host(k)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        step_host(k)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    h2d = W * H + 2 * NKP * 32 + ba_problem.n_cams * 57 + ba_problem.n_points * 25 + ba_problem.n_obs * (8 + 24)
    d2h = NKP * 60 + NKP * 12 + ba_problem.n_cams * 56 + ba_problem.n_points * 24

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    fps = world * args.steps / (ms * 1e-3)
    ab = algorithmic_bytes()
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    def gbs(nbytes, ms_):
        return nbytes / (ms_ * 1e-3) / 1e9
    ach = gbs(big_bytes, t_sweep_big)
    line = {"metric": "frames/sec detect+match+local-BA @1920x1080 mono", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
            "config": {"workload": workload, "l2": f"input ring of {RING} frames ({RING * W * H / 1e6:.0f} MB) > 126 MB L2",
                       "parallelism": f"replicas x{world}"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": world * e2e_steps / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps},
            "stages_ms": {"extract": t_ext, "match": t_match, "local_ba": t_ba},
            # dominant HBM-bound kernel of the path = the BA Jacobian sweep (BASELINE metric, 2nd clause), at config-5 size
            "roofline": {"kernel": "BA Jacobian sweep K6 (ba_linearize_kernel: camera pass + landmark pass in one launch), 500 cams/100k pts/1M obs",
                         "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": 150.6e6, "traffic_source": "profiles/r01_ncu_summary.md (dram__bytes_read 44.25 MB + dram__bytes_write 106.36 MB of one launch, ncu --set full)",
                         "peak_source": peak_src, "algorithmic_bytes": int(big_bytes), "ms_per_sweep": t_sweep_big},
            # the same quantity for the kernels as they run inside the timed step (one frame / one 10k-observation window:
            # launch-latency-bound, reported for completeness)
            "roofline_step": {"extract_chain": {"algorithmic_bytes": ab["extract"], "ms": t_ext, "achieved_gbs": gbs(ab["extract"], t_ext),
                                                "frac": gbs(ab["extract"], t_ext) / peak},
                              "ba_sweep_local": {"algorithmic_bytes": ab["ba_sweep"], "ms": t_sweep_local,
                                                 "achieved_gbs": gbs(ab["ba_sweep"], t_sweep_local), "frac": gbs(ab["ba_sweep"], t_sweep_local) / peak},
                              "match": {"algorithmic_bytes": ab["match"], "popc32": 8 * NKP * NKP, "ms": t_match,
                                        "gpopc_per_s": 8 * NKP * NKP / (t_match * 1e-3) / 1e9, "bound": "integer POPC pipe"}},
            }
    # CPU baseline on a bounded sample, rank 0 only, N=1 only
    if world == 1:
        try:
            n_cpu = 10
            cfps, what = cpu_path(base, ba_problem, n_cpu, cores)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"{n_cpu} frames: {what}"}
        except Exception as e:  # the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "error": repr(e)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
