#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
echo "== ncu full classify"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_classify -s 4 -c 2 -o gpurun_out/prof_classify_r1 -f $B > gpurun_out/ncu_full.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
echo "== variants"
V="python tools/kernel_variants.py"
( $V 10000000 10000 20000
  TB_UNROLL=4 $V 10000000 10000 20000
  TB_UNROLL=16 $V 10000000 10000 20000
  TB_TPB=64 $V 10000000 10000 20000
  TB_L2_FETCH_GRANULARITY=64 $V 10000000 10000 20000
  TB_L2_FETCH_GRANULARITY=128 $V 10000000 10000 20000
  $V 10000000 10000 20000 x 0.25
  $V 10000000 10000 20000 x 0.75
  $V 10000000 10000 20000 x 0.9
  $V 1000000 10000 20000
  $V 50000000 10000 20000
  $V 150000000 10000 20000
  $V 10000000 1000000 150 fixed
  K=31 $V 10000000 10000 20000
) > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log | grep -v "^Found\|^Done\|^Reading"
echo "== gather 64/128B blocks"
python - <<'PY' > gpurun_out/gather2.log 2>&1
import sys; sys.path.insert(0,'.')
from trio_binning_b200 import kmers
for mb in (320, 4096):
    for variant in (0, 3, 4):
        for mlp in (2, 4, 8):
            r, ms = kmers.gather_bench(mb << 20, mlp, 256, 8, 48, variant)
            span = {0:1,3:2,4:4}[variant]
            print(f"{mb}MB variant {variant} mlp {mlp}: {r/1e9:.2f} G units/s, {r*32*span/1e9:.0f} GB/s", flush=True)
PY
cat gpurun_out/gather2.log
