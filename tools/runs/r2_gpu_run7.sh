#!/bin/bash
# 2-GPU box: the product CLI on 1 and 2 GPUs (identical outputs), bench.py under torchrun
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_2gpu.txt
timeout 600 python tools/cli_multi_gpu.py 2 3000 8000 >> gpurun_out/r2_2gpu.txt 2>&1
timeout 900 python tools/cli_multi_gpu.py 2 40000 15000 >> gpurun_out/r2_2gpu.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native_pipeline or set_clone" 2>&1 | tail -3 ) >> gpurun_out/r2_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n2.json 2> gpurun_out/r2_bench_reference_n2.err
cat gpurun_out/r2_2gpu.txt
