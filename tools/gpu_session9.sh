#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== pipeline bench"; timeout 1500 python tools/pipeline_bench.py 10000 20000 > gpurun_out/pipeline.log 2>&1; cat gpurun_out/pipeline.log
