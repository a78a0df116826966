#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ba_gpu.py -m gpu -q -k "sweep or large_graph or bit_repro" > gpurun_out/pytest_new.log 2>&1; tail -3 gpurun_out/pytest_new.log
timeout 300 python tools/sweep_bench.py > gpurun_out/sweep_bench.log 2>&1; head -4 gpurun_out/sweep_bench.log
