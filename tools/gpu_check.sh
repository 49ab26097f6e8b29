#!/bin/bash
# What a round-end check runs on the GPU box: smoke, parity tests, both bench arms.
#   gpurun --timeout 1800 -- 'bash tools/gpu_check.sh'
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_reference.json | cut -c1-300
python bench.py 2>/dev/null | tee gpurun_out/bench_n1.json | cut -c1-400
