#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
echo "== CLI sharded check (N=2)"; timeout 600 python tools/cli_multi_gpu.py 2 > gpurun_out/cli_multi.log 2>&1; echo rc=$?; cat gpurun_out/cli_multi.log
for n in 8 4; do
echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 3 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err; echo "rc=$?"; tail -1 gpurun_out/bench_n$n.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value %.1f G' % (d['value']/1e9), 'e2e %.1f G' % (d['e2e']['value']/1e9), 'e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9), 'ms/step', d['ms_per_step'], 'clocks', d['clocks'])
"; tail -2 gpurun_out/bench_n$n.err
done
