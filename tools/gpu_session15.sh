#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -14 gpurun_out/topo.txt
echo "== bench N=8 with NUMA binding"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/bench_n8_numa.log 2> gpurun_out/bench_n8_numa.err; echo "rc=$?"; tail -1 gpurun_out/bench_n8_numa.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value %.1f G' % (d['value']/1e9), 'e2e %.1f G' % (d['e2e']['value']/1e9), 'e2e_packed %.1f G' % (d['e2e_packed']['value']/1e9), d['cpu_affinity'])
"; tail -2 gpurun_out/bench_n8_numa.err
