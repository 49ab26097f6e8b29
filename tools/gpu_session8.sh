#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','value_l2_flushed','ms_per_step'): print(k, d[k])
for k in ('e2e','e2e_packed','e2e_ascii_hostpack','host_pack'): print(k, {x:y for x,y in d[k].items() if x not in ('api','what')})
print('roofline frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'])
"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
from trio_binning_b200 import kmers, synth
ka, kb = synth.random_kmer_sets(1000, 1000, 21, 1)
bases, offsets, _ = synth.random_reads(10000, 20000, 21, ka, kb, seed=2)
total = int(offsets[-1])
pb = kmers.PackedBatch(total, len(offsets))
for th in (1, 2, 4, 8, 16, 24, 32):
    pb.pack(bases, offsets, 21, threads=th)
    t=time.perf_counter()
    for _ in range(5): pb.pack(bases, offsets, 21, threads=th)
    dt=(time.perf_counter()-t)/5
    print("host pack", th, "threads:", f"{total/dt/1e9:.2f} Gbases/s", flush=True)
PY
