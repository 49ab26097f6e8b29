#!/bin/bash
set -u
mkdir -p gpurun_out
for c in c2 c3 c4 c5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_classify_pair_mini -s 2 -c 1 -o gpurun_out/r2_${c}_mini -f python tools/config_bench.py $c --iters 2 > gpurun_out/r2_ncu_${c}.log 2>&1
  grep '"config"' gpurun_out/r2_ncu_${c}.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], d['windows'], round(d['classify_ms'],3))"
done
