#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 100 python profiles/microbench.py gt5 2>&1 | grep lstm > gpurun_out/lstm_time.txt
timeout 300 python bench.py --no-seq4000 > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/lstm_time.txt
python - <<'PY'
import json
for f in ('gpurun_out/bench_spat_gt5.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], d['dtype'], round(d['roofline']['ms_per_launch']*1e3,1))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
