#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/gq.py <<'PY'
import sys, os; sys.path.insert(0,'.')
from trio_binning_b200 import kmers
mb = int(sys.argv[1]); variants = [int(x) for x in sys.argv[2].split(',')]
for v in variants:
    r, ms = kmers.gather_bench(mb << 20, 8, 256, 8, 32, v)
    print(f"{mb}MB variant {v}: {r/1e9:.2f} G sectors/s  {ms:.3f} ms", flush=True)
PY
echo "== qualifiers, plain timing"
TB_VERBOSE=1 python /tmp/gq.py 4096 0,5,6,7,8,9,10,11,12 2>&1 | tee gpurun_out/gather_qual.log
TB_VERBOSE=1 TB_L2_FETCH_GRANULARITY=128 python /tmp/gq.py 4096 0 2>&1 | tee -a gpurun_out/gather_qual.log
TB_VERBOSE=1 TB_L2_FETCH_GRANULARITY=0 python /tmp/gq.py 4096 0 2>&1 | tee -a gpurun_out/gather_qual.log
echo "== qualifiers under ncu: dram bytes per sector request"
timeout 900 ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum --clock-control none -k regex:k_gather --csv --log-file gpurun_out/gather_qual_ncu.csv python /tmp/gq.py 4096 0,5,6,7,8,9,10,11,12,2 > /dev/null 2>&1
python - <<'PY' | tee -a gpurun_out/gather_qual.log
import csv
rows = list(csv.DictReader(l for l in open('gpurun_out/gather_qual_ncu.csv') if l.startswith('"')))
byid = {}
for r in rows:
    byid.setdefault(r['ID'], {'name': r['Kernel Name']})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
for i, d in byid.items():
    req = d.get('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 0)
    print(i, d['name'][:60], 'dramB/req=%.1f' % (d['dram__bytes_read.sum'] / max(req, 1)), 'ltsSect/req=%.2f' % (d['lts__t_sectors_srcunit_tex_op_read.sum'] / max(req, 1)), 'ms=%.3f' % (d['gpu__time_duration.sum'] / 1e6))
PY
echo "== L2 residency curve (variant 0)"
python - <<'PY' | tee gpurun_out/gather_l2curve.log
import sys; sys.path.insert(0,'.')
from trio_binning_b200 import kmers
for mb in (4, 8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 160, 200, 256, 320, 512, 1024):
    r, ms = kmers.gather_bench(mb << 20, 8, 256, 8, 64, 0)
    print(f"{mb}MB: {r/1e9:.2f} G sectors/s", flush=True)
PY
echo "== pcie"
python - <<'PY' | tee gpurun_out/pcie.log
import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
for name, fn in (("h2d", lambda: d.copy_(x, non_blocking=True)), ("d2h", lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * (1 << 30) / (time.perf_counter() - t) / 1e9))
PY
nvidia-smi -q | grep -A6 "GPU Link Info" | head -12 | tee -a gpurun_out/pcie.log
