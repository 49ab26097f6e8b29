#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 8 4 2; do
echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 50 --warmup 3 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err; echo "rc=$?"; tail -1 gpurun_out/bench_n$n.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value %.1f G' % (d['value']/1e9), 'e2e %.1f G via %s' % (d['e2e']['value']/1e9, d['e2e']['route']), 'ascii_device_pack %.1f G' % (d['e2e_variants']['ascii_device_pack']['value']/1e9), 'e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9), 'ms/step %.3f' % d['ms_per_step'])
"
done
