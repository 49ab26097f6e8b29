#!/bin/bash
# bench.py under torchrun at N = 2 with the h2d_ceiling object (plain H2D copies by all ranks at once)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "measurement_aids" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2_h2d.json 2> gpurun_out/r2_bench_n2_h2d.err
echo "bench N=2 rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2_h2d.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'n_gpus', 'h2d_ceiling')}, d['e2e']['value'])
PY
tail -3 gpurun_out/r2_bench_n2_h2d.err
