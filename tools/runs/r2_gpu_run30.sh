#!/bin/bash
# The pipeline leg of bench.py (file -> files) after the parallel gzip reader and the faster
# encoder; C2 sets to keep the set-up short.
mkdir -p gpurun_out
TB_PGZIP_DEBUG=1 timeout 200 python bench.py --workload c2 --steps 10 --no-cpu-baseline --no-gather --no-other-workloads \
    --pipeline-gbases 1.5 > gpurun_out/r2_pipeline_pgzip.json 2> gpurun_out/r2_pipeline_pgzip.err
echo "rc=$?"; grep pgzip gpurun_out/r2_pipeline_pgzip.err | head -3
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_pipeline_pgzip.json').read().strip().splitlines()[-1])
p = d['pipeline']
print(p.get('generate_s'), p.get('host_threads'))
for r in p['runs']:
    print(r['run'], round(r['wall_s'], 2), round(r['mbases_per_s']), {k: round(v, 2) for k, v in r['seconds'].items()})
PY
