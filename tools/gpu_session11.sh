#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== sanitizer (memcheck) on the small parity tests"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_counts or edge_shapes or invalid_bases or host_packed or prefilter or set_file or degenerate" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 gpurun_out/sanitizer.log
echo "== configs"
timeout 2400 python tools/run_configs.py c5 c3 c4 > gpurun_out/configs.log 2> gpurun_out/configs.err; echo "rc=$?"; cat gpurun_out/configs.log; tail -3 gpurun_out/configs.err
