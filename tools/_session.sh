#!/bin/bash
set -u
mkdir -p gpurun_out
V="python tools/kernel_variants.py"
run() { ( $V 10000000 10000 20000; TB_FILTER_UNROLL=32 $V 10000000 10000 20000; TB_FILTER_UNROLL=16 $V 10000000 10000 20000; $V 150000000 10000 20000; K=31 $V 400000000 10000 20000 ) 2>&1 | grep -v "^Found\|^Done\|^Reading" | sed -e 's/ lf=0.5 buckets=[0-9]*//' -e 's/UNROLL=[0-9]* TPB=128 L2G=32: //' -e 's/table_MB=[0-9x]* //' -e 's/(build.*//'; }
echo "== minblocks 5 (96 regs)"; run | tee gpurun_out/mb5.log
for mb in 6 8; do
  echo "== minblocks $mb"; TB_NVCC_EXTRA="-DTB_MINI_MIN_BLOCKS=$mb" python trio_binning_b200/_build.py --force > /dev/null 2>&1; run | tee gpurun_out/mb$mb.log
done
python trio_binning_b200/_build.py --force > /dev/null 2>&1
echo "== pytest gpu (default build)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
