#!/bin/bash
set -u
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f G  flushed %.1f G  e2e %.1f G  e2e_packed %.1f G (%.3f ms)  hostpack-e2e %.1f G  kernel %.3f ms  frac %.3f' % (d['value']/1e9, d['value_l2_flushed']/1e9, d['e2e']['value']/1e9, d['e2e_packed']['value']/1e9, d['e2e_packed']['ms_per_step'], d['e2e_ascii_hostpack']['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
"; }
B="python bench.py --no-cpu-baseline --steps 50"
echo "default";            $B > gpurun_out/b_default.log 2>/dev/null; show gpurun_out/b_default.log
echo "window+plain";       TB_L2_WINDOW=1 TB_FILTER_MODE=1 TB_VERBOSE=1 $B > gpurun_out/b_win1.log 2> gpurun_out/b_win1.err; show gpurun_out/b_win1.log; grep "L2 window" gpurun_out/b_win1.err | head -2
echo "window+evict_last";  TB_L2_WINDOW=1 TB_FILTER_MODE=2 $B > gpurun_out/b_win2.log 2>/dev/null; show gpurun_out/b_win2.log
echo "plain (normal)";     TB_FILTER_MODE=1 $B > gpurun_out/b_mode1.log 2>/dev/null; show gpurun_out/b_mode1.log
echo "filter 2^28";        TB_FILTER_LOG2=28 $B > gpurun_out/b_m28.log 2>/dev/null; show gpurun_out/b_m28.log
cp gpurun_out/b_default.log gpurun_out/bench_n1.log
