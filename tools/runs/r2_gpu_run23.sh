#!/bin/bash
# bitmap sizing rule (bits per key by k): parity + timings
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "c3 or c4 or c5 or k25 or random_batches or every_k or prefilter or config2 or short_read" 2>&1 | tail -3 ) > gpurun_out/r2_pytest_sizing.txt; cat gpurun_out/r2_pytest_sizing.txt
for c in c2 c3 c4 c5; do timeout 600 python tools/config_bench.py $c 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    x=json.loads(l); print(x['config'],'log2',x['filter_log2'],'pass',round(x['pass_fraction'],4),'ms',round(x['classify_ms'],3),round(x['gbases_per_s_kernel'],1))"; done
timeout 600 python tools/config_bench.py ksweep 2>/dev/null > gpurun_out/r2_ksweep_sized.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2_ksweep_sized.jsonl'):
    x=json.loads(l); print(x['config'], x['filter_log2'], round(x['pass_fraction'],4), round(x['classify_ms'],3), round(x['gbases_per_s_kernel'],1))
PY
