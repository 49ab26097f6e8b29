#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
V="python tools/kernel_variants.py"
( $V 10000000 10000 20000
  TB_FILTER_LOG2=0 $V 10000000 10000 20000
  TB_FILTER_LOG2=28 $V 10000000 10000 20000
  TB_FILTER_LOG2=30 $V 10000000 10000 20000
  TB_FILTER_UNROLL=8 $V 10000000 10000 20000
  TB_FILTER_UNROLL=32 $V 10000000 10000 20000
  TB_TPB=64 $V 10000000 10000 20000
  $V 1000000 10000 20000
  $V 50000000 10000 20000
  TB_FILTER_LOG2=0 $V 50000000 10000 20000
  $V 150000000 10000 20000
  TB_FILTER_LOG2=30 $V 150000000 10000 20000
  $V 10000000 1000000 150 fixed
  K=31 $V 10000000 10000 20000
) > gpurun_out/variants4.log 2>&1
grep -v "^Found\|^Done\|^Reading" gpurun_out/variants4.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4.log 2> gpurun_out/bench4.err; echo "bench rc=$?"; tail -2 gpurun_out/bench4.log; tail -3 gpurun_out/bench4.err
