#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather"
for m in 28 29; do
TB_FILTER_LOG2=$m timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_classify_filtered -s 4 -c 1 -o gpurun_out/prof_filtered_m$m -f $B > gpurun_out/ncu_f$m.log 2>&1; echo rc=$?
done
ls -la gpurun_out/*.ncu-rep
