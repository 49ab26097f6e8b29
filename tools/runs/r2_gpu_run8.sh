#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | grep -v "^Found\|^Done\|^Reading" | tail -14 ) > gpurun_out/r2_pytest_gpu8.txt 2>&1
cat gpurun_out/r2_pytest_gpu8.txt
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python __graft_entry__.py --smoke 2>&1 | tail -2
