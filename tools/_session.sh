#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench N=1"; timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f G' % (d['value']/1e9), 'flushed %.1f G' % (d['value_l2_flushed']/1e9), 'primary e2e %.1f G via %s' % (d['e2e']['value']/1e9, d['e2e']['route']), 'e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9))
r=d['roofline']; print(r['kernel'], 'kernel_ms %.3f' % r['kernel_ms'], 'frac %.3f' % r['frac'], 'share', r['kernel_share_of_step'])
"
echo "== configs"; timeout 2400 python tools/run_configs.py c5 c3 c4 > gpurun_out/configs.log 2> gpurun_out/configs.err; python - <<'PY'
import json
for l in open('gpurun_out/configs.log'):
    d=json.loads(l); print(d['config'], 'filter 2^%d bits=%d pass=%.4f classify %.3f ms  %.1f Gbases/s kernel  %.1f step  parity mismatches %s' % (d['filter_log2'], d['bits_per_key'], d['filter_pass_fraction'], d['classify_ms'], d['gbases_per_s_kernel'], d['gbases_per_s_step'], d['parity']['mismatches']))
PY
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gather"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_classify_pair_mini -s 6 -c 1 -o gpurun_out/prof_classify_mini_r1 -f $B > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_mini.csv $B > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
