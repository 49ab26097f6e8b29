#!/bin/bash
# After the packed counting slots: the whole GPU suite, then the counter's throughput from pageable
# and from page-locked memory (k = 21 packed, k = 25 split).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for k in 21 25; do
    timeout 600 python tools/count_bench.py --k $k > gpurun_out/r2_count_bench_k$k.json 2> gpurun_out/r2_count_bench_k$k.err
    echo "count_bench k=$k rc=$?"; tail -c 1500 gpurun_out/r2_count_bench_k$k.json
done
