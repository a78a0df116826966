#!/bin/bash
# 8-GPU box: config-5 global BA (strong scaling) on 1 / 2 / 4 / 8 GPUs, the plugin-level multi-GPU test, the vocabulary plugin test
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n8.txt
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) tools/global_ba_bench.py > gpurun_out/gba_scale_n$n.json 2> gpurun_out/gba_scale_n$n.err
  cat gpurun_out/gba_scale_n$n.json
done
timeout 600 python -m pytest tests/test_plugins.py tests/test_dist.py -m gpu -q -k "sharded or vocabulary or nccl or world" > gpurun_out/pytest_n8.log 2>&1; tail -4 gpurun_out/pytest_n8.log
