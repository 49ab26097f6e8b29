#!/bin/bash
# round-2 first GPU call: parity tests incl. full-size configs, config timings, bitmap-load variants at C3,
# k sweep, hit-density sweep, ncu capture of the classify kernel at C3
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpuinfo.txt 2>&1
free -g >> gpurun_out/r2_gpuinfo.txt; nproc >> gpurun_out/r2_gpuinfo.txt
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -s 2>&1 | grep -v "^Found\|^Done\|^Reading" | tail -40 ) > gpurun_out/r2_pytest_gpu.txt 2>&1
for c in c2 c3 c4 c5; do timeout 600 python tools/config_bench.py $c 2>/dev/null; done > gpurun_out/r2_configs_base.jsonl
KS=21 SHAPE="150000000 10000 25000" timeout 900 bash tools/variants.sh run > gpurun_out/r2_variants_bitmap_ld_c3.txt 2>&1
timeout 600 python tools/config_bench.py ksweep 2>/dev/null > gpurun_out/r2_ksweep.jsonl
timeout 600 python tools/config_bench.py density 2>/dev/null > gpurun_out/r2_density.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_classify_pair_mini -s 2 -c 1 -o gpurun_out/r2_c3_mini python tools/config_bench.py c3 --iters 2 > gpurun_out/r2_ncu_c3.log 2>&1
ls -la gpurun_out
