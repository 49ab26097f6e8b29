#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_pair.csv $B > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
echo "== ncu full pair kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_classify_pair -s 6 -c 1 -o gpurun_out/prof_classify_pair_r1 -f $B > gpurun_out/ncu_full.log 2>&1; echo rc=$?
echo "== ncu full pack kernel"
timeout 1200 ncu --set full --clock-control none -k regex:k_pack -s 6 -c 1 -o gpurun_out/prof_pack_r1 -f $B > gpurun_out/ncu_pack.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
