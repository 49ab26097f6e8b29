#!/bin/bash
set -u
mkdir -p gpurun_out
grep -m1 "model name" /proc/cpuinfo; grep -m1 -o "avx512bw" /proc/cpuinfo | head -1; nproc
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f G' % (d['value']/1e9), 'primary e2e %.1f G via %s' % (d['e2e']['value']/1e9, d['e2e']['route']))
for k,v in d['e2e_variants'].items(): print('  ', k, '%.1f G' % (v['value']/1e9), '%.3f ms' % v['ms_per_step'])
print('e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9), 'host_pack %.1f G on %d threads' % (d['host_pack']['value']/1e9, d['host_pack']['threads']))
"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
from trio_binning_b200 import kmers, synth
ka, kb = synth.random_kmer_sets(1000, 1000, 21, 1)
bases, offsets, _ = synth.random_reads(10000, 20000, 21, ka, kb, seed=2)
total = int(offsets[-1])
pb = kmers.PackedBatch(total, len(offsets))
for th in (1, 2, 4, 8, 16):
    pb.pack(bases, offsets, 21, threads=th)
    t=time.perf_counter()
    for _ in range(5): pb.pack(bases, offsets, 21, threads=th)
    dt=(time.perf_counter()-t)/5
    print("host pack (avx512 if present)", th, "threads:", f"{total/dt/1e9:.2f} Gbases/s", flush=True)
PY
