#!/bin/bash
# sanitizers on the small parity tests + the new kernels; 5-Gbase pipeline leg
set -u
mkdir -p gpurun_out
df -h /tmp | tail -1 > gpurun_out/r2_df.txt
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_find_unique_kmers.py -m gpu -x -q -k "golden_counts or edge_shapes or invalid_bases or kmer_in_hash_set or every_k_vs_oracle[13] or every_k_vs_oracle[25] or every_k_vs_oracle[32] or set_clone or counting_histogram_and_subtraction_vs_oracle[21] or partitions" 2>&1 | tail -15 ) > gpurun_out/r2_sanitizer_memcheck.txt
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_find_unique_kmers.py -m gpu -x -q -k "golden_counts or set_file_quirks or every_k_vs_oracle[21] or counting_histogram_and_subtraction_vs_oracle[21]" 2>&1 | tail -15 ) > gpurun_out/r2_sanitizer_racecheck.txt
tail -3 gpurun_out/r2_sanitizer_memcheck.txt gpurun_out/r2_sanitizer_racecheck.txt
timeout 1500 python bench.py --workload c2 --steps 10 --no-cpu-baseline --no-other-workloads --no-gather --pipeline-gbases 5 > gpurun_out/r2_pipeline_5gbases.json 2> gpurun_out/r2_pipeline_5gbases.err
tail -3 gpurun_out/r2_pipeline_5gbases.err
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err
