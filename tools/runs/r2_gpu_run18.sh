#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "short_read or c5 or c3 or k25 or edge_shapes or batch_shape or random_batches or invalid_bases or golden or long_runs or every_k or two_level" 2>&1 | tail -3 ) > gpurun_out/r2_pytest_bounds.txt; cat gpurun_out/r2_pytest_bounds.txt
echo "== C2"; KS="21 31" SHAPE="10000000 10000 20000" timeout 600 bash tools/variants.sh run
echo "== C3"; KS="21" SHAPE="150000000 10000 25000" timeout 600 bash tools/variants.sh run
echo "== C5"; KS="21" SHAPE="50000000 2000000 150 fixed" timeout 600 bash tools/variants.sh run
