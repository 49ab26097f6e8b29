#!/bin/bash
# 8-GPU box: bench.py under torchrun at N=8 (what the driver's scaling run does), the product CLI on 8 GPUs
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_8gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -2 gpurun_out/r2_bench_n8.err | cut -c1-300
timeout 400 python tools/cli_multi_gpu.py 8 40000 15000 >> gpurun_out/r2_8gpu.txt 2>&1
cat gpurun_out/r2_8gpu.txt | cut -c1-400
