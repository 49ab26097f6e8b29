#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "two_level or c3 or c5 or every_k_vs_oracle[2" 2>&1 | grep -v "^Found\|^Done\|^Reading" | tail -6 ) > gpurun_out/r2_pytest_hier.txt
cat gpurun_out/r2_pytest_hier.txt
for c in c3 c5 c2; do for h in 1 0; do TB_HIER=$h timeout 600 python tools/config_bench.py $c 2>/dev/null; done; done > gpurun_out/r2_configs_hier.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r2_configs_hier.jsonl'):
    x=json.loads(l); print(x['config'], x['env'], 'pass', round(x['pass_fraction'],4), 'ms', round(x['classify_ms'],3), round(x['gbases_per_s_kernel'],1))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_classify_pair_mini -s 2 -c 1 -o gpurun_out/r2_c3_hier -f python tools/config_bench.py c3 --iters 2 > gpurun_out/r2_ncu_c3_hier.log 2>&1
