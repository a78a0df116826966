#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_sweep_kernel -c 1 -o gpurun_out/r02_sweep_final python tools/sweep_ncu.py > gpurun_out/ncu_sweep.log 2>&1; tail -2 gpurun_out/ncu_sweep.log
