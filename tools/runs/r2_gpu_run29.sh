#!/bin/bash
# bench.py under torchrun at N = 8 with the h2d_ceiling object
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 \
    bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8_h2d.json 2> gpurun_out/r2_bench_n8_h2d.err
echo "bench N=8 rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n8_h2d.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'h2d_ceiling')}, d['e2e']['value'], {k: v['value'] for k, v in d['e2e_variants'].items()})
PY
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2_topo_8gpu.txt
