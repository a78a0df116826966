#!/bin/bash
# End-of-round validation on one GPU: gpu test suite, default bench line, reference arm, ncu launch lists (bench step, global BA), sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 3 --warmup 3 --no-global-ba > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/globalba_launches.csv python tools/global_ba_bench.py --reps 1 > gpurun_out/gba_under_ncu.log 2>&1; echo "ncu gba rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 300 python tools/bow_bench.py > gpurun_out/bow_bench.log 2>&1; cat gpurun_out/bow_bench.log
cat gpurun_out/bench_n1.json | cut -c1-300
