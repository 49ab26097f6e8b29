#!/bin/bash
set -u
mkdir -p gpurun_out
V="python tools/kernel_variants.py"
( for mode in 0 1 2; do for m in 28 29; do
    TB_FILTER_MODE=$mode TB_FILTER_LOG2=$m $V 10000000 10000 20000
  done; done
  TB_FILTER_MODE=2 TB_L2_PERSIST_MB=0 TB_FILTER_LOG2=28 $V 10000000 10000 20000
  TB_FILTER_MODE=2 TB_L2_PERSIST_MB=0 TB_FILTER_LOG2=29 $V 10000000 10000 20000
  TB_FILTER_MODE=2 TB_FILTER_LOG2=27 $V 10000000 10000 20000
  TB_FILTER_MODE=2 TB_FILTER_LOG2=30 $V 10000000 10000 20000
  TB_FILTER_MODE=1 TB_FILTER_LOG2=30 $V 10000000 10000 20000
  TB_FILTER_MODE=2 $V 50000000 10000 20000
  TB_FILTER_MODE=2 $V 150000000 10000 20000
  TB_FILTER_MODE=2 TB_FILTER_LOG2=30 $V 150000000 10000 20000
  TB_FILTER_MODE=1 TB_FILTER_LOG2=30 $V 150000000 10000 20000
) > gpurun_out/variants6.log 2>&1
grep -v "^Found\|^Done\|^Reading" gpurun_out/variants6.log | sed -e 's/ lf=0.5 buckets=[0-9]*//' -e 's/UNROLL=8 TPB=128 L2G=32: //'
TB_FILTER_MODE=2 TB_FILTER_LOG2=28 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_classify_filtered -s 4 -c 1 -o gpurun_out/prof_filtered_m28_keep -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather > gpurun_out/ncu_f28k.log 2>&1; echo rc=$?
