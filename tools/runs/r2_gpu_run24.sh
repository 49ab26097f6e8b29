#!/bin/bash
# find-unique-kmers compute: throughput (no profiler), then launch times and one full capture of
# k_count_windows on a smaller instance of the same workload.
mkdir -p gpurun_out
timeout 600 python tools/count_bench.py > gpurun_out/r2_count_bench.json 2> gpurun_out/r2_count_bench.err
echo "count_bench rc=$?"; tail -c 1500 gpurun_out/r2_count_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_count|k_pack' -c 60 --csv \
    --log-file gpurun_out/r2_launches_count.csv python tools/count_bench.py --genome 5000000 --cpu-sample 300000 \
    > gpurun_out/r2_count_ncu_times.log 2>&1
echo "ncu times rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_count_windows -s 2 -c 1 -f \
    -o gpurun_out/r2_count_windows python tools/count_bench.py --genome 5000000 --cpu-sample 300000 \
    > gpurun_out/r2_count_ncu_full.log 2>&1
echo "ncu full rc=$?"
