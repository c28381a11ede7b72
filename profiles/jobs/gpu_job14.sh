#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-seq4000 > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 300 python bench.py --workload spat_p100 --steps 20 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
tail -5 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ('gpurun_out/bench_spat_gt5.json','gpurun_out/bench_spat_p100.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
