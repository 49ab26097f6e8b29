#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== C2 shape (10M+10M, 10k x 20kb)" > gpurun_out/r2_variants_alu.txt
KS="21 31" SHAPE="10000000 10000 20000" timeout 900 bash tools/variants.sh run >> gpurun_out/r2_variants_alu.txt 2>&1
echo "== C3 shape (150M+150M, 10k x 25kb)" >> gpurun_out/r2_variants_alu.txt
KS="21" SHAPE="150000000 10000 25000" timeout 900 bash tools/variants.sh run >> gpurun_out/r2_variants_alu.txt 2>&1
echo "== C5 shape (50M+50M, 2M x 150)" >> gpurun_out/r2_variants_alu.txt
KS="21" SHAPE="50000000 2000000 150 fixed" timeout 900 bash tools/variants.sh run >> gpurun_out/r2_variants_alu.txt 2>&1
timeout 600 python tools/config_bench.py ksweep > gpurun_out/r2_ksweep.jsonl 2> gpurun_out/r2_ksweep.err
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_k or random_batches" 2>&1 | tail -3 ) > gpurun_out/r2_pytest_k.txt
cat gpurun_out/r2_variants_alu.txt
