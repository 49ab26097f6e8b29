#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log; tail -2 gpurun_out/bench_n1.err
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log; tail -2 gpurun_out/bench_ref.err
