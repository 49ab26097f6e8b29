#!/bin/bash
# Packed counting slots (k <= 22): parity tests, throughput, launch times, one full capture,
# memcheck on the counter tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_find_unique_kmers.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/count_bench.py > gpurun_out/r2_count_bench_packed.json 2> gpurun_out/r2_count_bench_packed.err
echo "count_bench k=21 rc=$?"; tail -c 1200 gpurun_out/r2_count_bench_packed.json
timeout 600 python tools/count_bench.py --k 25 > gpurun_out/r2_count_bench_k25.json 2>> gpurun_out/r2_count_bench_packed.err
echo "count_bench k=25 rc=$?"; tail -c 1200 gpurun_out/r2_count_bench_k25.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_count|k_pack' -c 60 --csv \
    --log-file gpurun_out/r2_launches_count_packed.csv python tools/count_bench.py --genome 5000000 --cpu-sample 300000 \
    > gpurun_out/r2_count_ncu_times.log 2>&1
echo "ncu times rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_count_windows -s 2 -c 1 -f \
    -o gpurun_out/r2_count_windows_packed python tools/count_bench.py --genome 5000000 --cpu-sample 300000 \
    > gpurun_out/r2_count_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_find_unique_kmers.py -m gpu -x -q \
    -k "saturated or partitions" > gpurun_out/r2_compute_sanitizer_memcheck_counter.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_compute_sanitizer_memcheck_counter.txt
