#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | grep -v "^Found\|^Done\|^Reading" | tail -30 ) > gpurun_out/r2_pytest_gpu4.txt 2>&1
tail -4 gpurun_out/r2_pytest_gpu4.txt
timeout 600 python tools/config_bench.py density > gpurun_out/r2_density.jsonl 2> gpurun_out/r2_density.err; tail -2 gpurun_out/r2_density.err
for c in c2 c3 c4 c5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_classify_pair_mini -s 2 -c 1 -o gpurun_out/r2_${c}_mini -f python tools/config_bench.py $c --iters 2 > gpurun_out/r2_ncu_${c}.log 2>&1
  grep '"config"' gpurun_out/r2_ncu_${c}.log | cut -c1-300
done
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-pipeline --no-other-workloads --no-gather > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out | head -40
