"""Writes profiles/r02_ncu_summary.md from the committed artifacts of the end-of-round validation run (tools/gpu_round_final.sh):
r02_bench_n1.json, r02_bench_ref.json, r02_bench_launches.csv, r02_globalba_launches.csv, r02_sweep_kernel.ncu-rep."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
d = json.load(open(os.path.join(P, "r02_bench_n1.json"))); r = json.load(open(os.path.join(P, "r02_bench_ref.json")))
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True, cwd=ROOT).stdout
out = []
w = out.append
w("# Round 2 — measured evidence (B200, sm_100a, clocks 1965 MHz, no throttle reasons)\n")
w("All numbers come from this repository's own commands on a B200 box (`tools/gpu_round_final.sh` under `gpurun`); none was taken under a profiler\n"
  "unless it says ncu.  Files: `r02_bench_n1.json` (the default `python bench.py` line), `r02_bench_ref.json` (`--impl reference`),\n"
  "`r02_bench_launches.csv` / `r02_globalba_launches.csv` (ncu launch lists), `r02_pytest_gpu.log` (+ `_n2`, `_n8`), `r02_sweep_bench.log`,\n"
  "`r02_sweep_kernel.ncu-rep` (+ `r02_sweep_traffic.json`), `r02_bow_bench.log`, `r02_sass_evidence.md`, `r02_global_ba_scaling.md`, `r02_sanitizer.md`.\n")
w("## 1. The step (1920x1080, 2000 kp extract + 2000x2000 match + local BA 50 KF / 2000 pts / 10 000 obs, 10 LM x 50 PCG)\n")
w("| quantity | value |\n|---|---|")
w(f"| `value` (device-resident, tracking and mapping pipelined on one GPU) | {d['value']:.1f} frames/s ({d['ms_per_step']:.3f} ms per step) |")
w(f"| `serial` (one stream, round-1 definition) | {d['serial']['value']:.1f} frames/s |")
e = d['e2e']
w(f"| `e2e` (host-buffer C-ABI, pageable frames, H2D {e['h2d_bytes_per_step']} B + D2H {e['d2h_bytes_per_step']} B per step in the timed region) | {e['value']:.1f} frames/s (pinned {e['pinned']:.1f}; serial pageable {e['serial_pageable']:.1f}) |")
w(f"| stages (ms) | extract {d['stages_ms']['extract']:.3f}, match {d['stages_ms']['match']:.3f}, local BA {d['stages_ms']['local_ba']:.3f} |")
w(f"| CPU arm on the same box: {r['cpu_baseline']['sample'].split(': ',1)[-1]} | {r['value']:.1f} frames/s |")
w(f"| config 4 (stereo 752x480: 2 x extract + stereo match + temporal match + PnP-RANSAC + local BA) | {d['config4_stereo']['value']:.1f} frames/s |")
g = d['global_ba']
w(f"| global BA, config 5, one GPU | {g['ms_per_lm_iteration']:.3f} ms per LM iteration (cost {g['initial_cost']:.2f} -> {g['cost']:.4f}) |")
b = d['bow_transform']
w(f"| BoW transform, 2000 descriptors, k=10 L=5 | {b['us_per_transform_host_buffers']:.0f} us through Python with host buffers, {b['us_per_transform_device_resident_descriptors']:.0f} us chained after the extraction, vs {b['cpu_reference_us']:.0f} us for the reference class (1 thread; published 615.5 us); parity {b['parity']} |")
rf = d['roofline']
w(f"| roofline (BA sweep at config 5, not in the timed step) | {rf['ms_per_sweep']*1e3:.1f} us per sweep = {rf['achieved']:.0f} GB/s algorithmic = {rf['frac']:.3f} of {rf['peak']} GB/s; on measured DRAM bytes {rf['frac_on_measured_dram_bytes']:.3f} |")
m = d['roofline_step']['match']
w(f"| match vs the measured POPC peak | {m['gpopc_per_s']:.0f} of {m['peak_gpopc_per_s']:.0f} Gpopc/s = {m['frac']:.2f} (launch sized for capacities, clipped by the device-side count) |")
w("\n## 2. Share of the step by kernel (ncu launch list of `bench.py --steps 3 --warmup 3 --no-global-ba`: cold-cache, serialised — the SHARES carry over)\n")
w(run("tools/summarize_launches.py", "profiles/r02_bench_launches.csv"))
w("## 3. One global-BA solve at config 5 (ncu launch list of `tools/global_ba_bench.py --reps 1`: two solves of 5 LM iterations)\n")
w(run("tools/summarize_launches.py", "profiles/r02_globalba_launches.csv"))
w("## 4. The BA sweep at config 5 (`ba_sweep_kernel`, one `ncu --set full` launch, `r02_sweep_kernel.ncu-rep`)\n")
w(run("tools/sweep_ncu_table.py"))
w("""Reading: 148 CTAs x 512 threads (16 warps/SM at 126 registers), 216 KB of dynamic shared memory per CTA.  Issue slots 35 % busy, fp64
pipe 24 %, tensor pipe 0 (nothing here is a GEMM: K = 2-3 contractions per observation, see DESIGN.md section 4).  DRAM traffic 145.4 MB
per launch against 177.7 MB algorithmic (part of the 144 MB of W blocks is still in the 126 MB L2 when the kernel ends) -> no wasted
traffic.  Stall profile per issued instruction: short scoreboard 2.5 + MIO throttle 1.5 (shared memory: pose-table gathers of 32
different camera rows per warp, term-major contribution tile, staged item data; 7.2 M bank conflicts -- 7.7 M before the rows were padded), long scoreboard 2.0 (was 4.7 before the cp.async stage
prefetch and the spill removal), wait 2.0 (dependent fp64 chains), barrier 0.7 (was 2.6 before the per-landmark V^-1 left the team's
critical path).  History of the kernel on this graph (`tools/sweep_bench.py`, CUDA events, 20 launches): ba.cu kernel 106.4 us ->
persistent 2x256-thread CTAs + ticket 102.5 -> 4 teams + host plan + cp.async stages 106.6 -> sums written by their threads, V^-1 moved
to the Schur preparation 95.5 -> camera pose re-read from shared memory (no spills), tree sums 88.4 us -> pose-table rows padded to an
odd number of 16-byte chunks 86.3 us (2.06 TB/s, 31.8 % of 6484.6 GB/s).

## 5. SASS evidence

`r02_sass_evidence.md`: `UTMALDG.2D` + `SYNCS` (mbarrier) in `orb_fast_kernel` (TMA tile loads), `UBLKCP` (bulk async copy, TMA engine)
+ `LDGSTS` (cp.async) in `ba_sweep_kernel`, `LDGSTS` in `ba_schur_chunks_kernel`, `UCGABAR` (hardware cluster barriers) in the
cluster PCG kernels.  No `HMMA/DMMA/UTCMMA`: no tensor-pipe instruction anywhere, by design (DESIGN.md section 4).
""")
open(os.path.join(P, "r02_ncu_summary.md"), "w").write("\n".join(out))
print("wrote profiles/r02_ncu_summary.md")
