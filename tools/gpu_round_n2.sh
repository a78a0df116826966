#!/bin/bash
# 2-GPU box: the whole gpu suite (NCCL tests included), then the config-5 global BA on 1 and 2 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n2.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_n2.log
tail -6 gpurun_out/pytest_gpu_n2.log
timeout 300 python tools/global_ba_bench.py > gpurun_out/gba_n1.json 2> gpurun_out/gba_n1.err; cat gpurun_out/gba_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/global_ba_bench.py > gpurun_out/gba_n2.json 2> gpurun_out/gba_n2.err; cat gpurun_out/gba_n2.json
