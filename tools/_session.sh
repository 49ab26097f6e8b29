#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f G' % (d['value']/1e9), 'flushed %.1f G' % (d['value_l2_flushed']/1e9), 'ms/step %.3f' % d['ms_per_step'], 'primary e2e %.1f G via %s' % (d['e2e']['value']/1e9, d['e2e']['route']), 'e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9))
r=d['roofline']; print(r['kernel'], 'kernel_ms %.3f' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'share', r['kernel_share_of_step'])
"
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>/dev/null; tail -1 gpurun_out/bench_ref.log | cut -c1-160
echo "== configs"; timeout 2400 python tools/run_configs.py c5 c3 c4 > gpurun_out/configs.log 2> gpurun_out/configs.err; python - <<'PY'
import json
for l in open('gpurun_out/configs.log'):
    d=json.loads(l); print(d['config'], 'filter 2^%d bits=%d pass=%.4f classify %.3f ms  %.1f Gbases/s kernel  %.1f step  parity mismatches %s' % (d['filter_log2'], d['bits_per_key'], d['filter_pass_fraction'], d['classify_ms'], d['gbases_per_s_kernel'], d['gbases_per_s_step'], d['parity']['mismatches']))
PY
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_mini.csv $B > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
