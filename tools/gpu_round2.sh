#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_bow_gpu.py -m gpu -x -q > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
timeout 300 python tools/sweep_bench.py > gpurun_out/sweep_bench.log 2>&1; cat gpurun_out/sweep_bench.log
timeout 300 python tools/bow_bench.py > gpurun_out/bow_bench.log 2>&1; cat gpurun_out/bow_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_sweep_kernel -c 2 -o gpurun_out/r02_sweep2 python tools/sweep_ncu.py > gpurun_out/ncu_sweep.log 2>&1; tail -3 gpurun_out/ncu_sweep.log
