#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_posegraph_gpu.py tests/test_dist.py -m gpu -q > gpurun_out/pytest_repro.log 2>&1; tail -4 gpurun_out/pytest_repro.log
timeout 300 python tools/global_ba_bench.py > gpurun_out/gba_n1.json 2> gpurun_out/gba_n1.err; cat gpurun_out/gba_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/globalba_launches.csv python tools/global_ba_bench.py --reps 1 > gpurun_out/gba_under_ncu.log 2>&1; echo "ncu gba rc=$?"
