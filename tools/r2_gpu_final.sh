#!/bin/bash
# what the driver runs at round end, plus the launch list of the bench command
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2_pytest_gpu_final.txt 2>&1; cat gpurun_out/r2_pytest_gpu_final.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-pipeline --no-other-workloads --no-gather > gpurun_out/r2_launches_bench.log 2>&1
