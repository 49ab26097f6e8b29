#!/bin/bash
set -u
mkdir -p gpurun_out
LIB=$(ls trio_binning_b200/kmers_c*.so | head -1)
cp $LIB /tmp/tb_orig.so
for v in hb4 hb6 hb7; do
  cp _variants/$v.so $LIB
  for spec in "c3 1 0" "c3 1 34" "c5 1 0" "c5 1 32" "c3 0 0" "c5 0 0"; do
    set -- $spec
    if [ "$3" = "0" ]; then unset TB_FILTER_LOG2; else export TB_FILTER_LOG2=$3; fi
    echo -n "$v $1 hier=$2 log2=$3: "
    TB_HIER=$2 timeout 600 python tools/config_bench.py $1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    x=json.loads(l); print('log2',x['filter_log2'],'pass',round(x['pass_fraction'],4),'ms',round(x['classify_ms'],3),round(x['gbases_per_s_kernel'],1))"
  done
done > gpurun_out/r2_variants_hier.txt 2>&1
cp /tmp/tb_orig.so $LIB
cat gpurun_out/r2_variants_hier.txt
