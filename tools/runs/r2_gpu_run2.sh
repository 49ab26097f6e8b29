#!/bin/bash
# round-2 second GPU call: new pipeline tests, new bench.py (both arms), ksweep, density (fixed)
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | grep -v "^Found\|^Done\|^Reading" | tail -30 ) > gpurun_out/r2_pytest_gpu2.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -5 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -3 gpurun_out/r2_bench_reference.err
timeout 600 python tools/config_bench.py ksweep > gpurun_out/r2_ksweep.jsonl 2> gpurun_out/r2_ksweep.err
tail -3 gpurun_out/r2_ksweep.err
ls -la gpurun_out
