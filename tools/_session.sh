#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench N=1"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f G' % (d['value']/1e9), 'ms/step %.3f' % d['ms_per_step'], 'primary e2e %.1f G via %s' % (d['e2e']['value']/1e9, d['e2e']['route']), 'e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9))
r=d['roofline']; print(r['kernel'], 'kernel_ms %.3f' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'share', r['kernel_share_of_step'])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pack -c 6 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather 2>/dev/null | grep k_pack | tail -3 | cut -d, -f5,15
