#!/bin/bash
mkdir -p gpurun_out
export GB_DEBUG=1
timeout 900 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "sweep or large_graph" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -5 gpurun_out/pytest_new.log
timeout 600 python -m pytest tests/test_plugins.py -m gpu -x -q -k "vocabulary" -s > gpurun_out/pytest_voc.log 2>&1; grep -n "bow:\|gslam_b200 voc\|passed\|failed" gpurun_out/pytest_voc.log | head -12
timeout 300 python tools/sweep_bench.py > gpurun_out/sweep_bench.log 2>&1; head -4 gpurun_out/sweep_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_sweep_kernel -c 1 -o gpurun_out/r02_sweep6 python tools/sweep_ncu.py > gpurun_out/ncu_sweep.log 2>&1; tail -2 gpurun_out/ncu_sweep.log
