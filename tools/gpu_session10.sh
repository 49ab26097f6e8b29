#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
V="python tools/kernel_variants.py"
( $V 10000000 10000 20000
  TB_FILTER_BITS=1 $V 10000000 10000 20000
  TB_PAIR=0 $V 10000000 10000 20000
  TB_FILTER_LOG2=28 $V 10000000 10000 20000
  $V 1000000 10000 20000
  $V 50000000 10000 20000
  TB_FILTER_BITS=2 $V 50000000 10000 20000
  TB_PAIR=0 $V 50000000 10000 20000
  $V 150000000 10000 20000
  TB_FILTER_BITS=2 $V 150000000 10000 20000
  TB_FILTER_LOG2=29 $V 150000000 10000 20000
  TB_PAIR=0 $V 150000000 10000 20000
  $V 10000000 1000000 150 fixed
  K=31 $V 10000000 10000 20000
  K=31 $V 150000000 10000 20000
) > gpurun_out/variants10.log 2>&1
grep -v "^Found\|^Done\|^Reading" gpurun_out/variants10.log | sed -e 's/ lf=0.5 buckets=[0-9]*//' -e 's/UNROLL=[0-9]* TPB=128 L2G=32: //' -e 's/FU=16 //'
