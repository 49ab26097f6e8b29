#!/bin/bash
# First GPU session: smoke, parity tests, bench, gather sweep, ncu launch list.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== gather sweep"; timeout 600 python tools/gather_sweep.py > gpurun_out/gather.log 2>&1; tail -40 gpurun_out/gather.log
