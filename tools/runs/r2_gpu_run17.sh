#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "short_read or c5 or c3 or edge_shapes or batch_shape or random_batches or invalid_bases or golden or long_runs or every_k_vs_oracle[21]" 2>&1 | tail -3 ) > gpurun_out/r2_pytest_vmask.txt; cat gpurun_out/r2_pytest_vmask.txt
for c in c5 c2 c3; do timeout 600 python tools/config_bench.py $c 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    x=json.loads(l); print(x['config'],'ms',round(x['classify_ms'],3),round(x['gbases_per_s_kernel'],1))"; done
