#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_2gpu.txt
timeout 300 python tools/cli_multi_gpu.py 2 3000 8000 >> gpurun_out/r2_2gpu.txt 2>&1
timeout 600 python tools/cli_multi_gpu.py 2 40000 15000 >> gpurun_out/r2_2gpu.txt 2>&1
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "native_pipeline or set_clone" 2>&1 | tail -3 ) >> gpurun_out/r2_2gpu.txt
cat gpurun_out/r2_2gpu.txt
