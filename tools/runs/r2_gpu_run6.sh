#!/bin/bash
set -u
mkdir -p gpurun_out
( TB_P2_WIDTH=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q -k "random_batches or every_k_vs_oracle[21] or every_k_vs_oracle[31] or invalid_bases or golden or c3 or c5 or canonical or short_read" 2>&1 | tail -4 ) > gpurun_out/r2_pytest_coop.txt
cat gpurun_out/r2_pytest_coop.txt
for w in 2 1 4; do TB_P2_WIDTH=$w timeout 600 python tools/config_bench.py density 2>/dev/null > gpurun_out/r2_density_p2_$w.jsonl; done
for c in c2 c3 c5; do for w in 2 1; do TB_P2_WIDTH=$w timeout 600 python tools/config_bench.py $c 2>/dev/null; done; done > gpurun_out/r2_configs_p2.jsonl
timeout 1500 python bench.py --workload c2 --steps 10 --no-cpu-baseline --no-other-workloads --no-gather --pipeline-gbases 5 > gpurun_out/r2_pipeline_5gbases.json 2> gpurun_out/r2_pipeline_5gbases.err
tail -2 gpurun_out/r2_pipeline_5gbases.err
