#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_model_tc.py -x -q > gpurun_out/pytest_part.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_part.log
tail -4 gpurun_out/pytest_part.log
for s in 0 16 32 48; do
  VOG_LANG_SMS=$s VOG_LANG_SPLIT_MIN_P=$([ $s -eq 0 ] && echo 1000000 || echo 2000) timeout 300 python bench.py --workload spat_p100 --steps 30 --no-seq4000 --no-cpu-baseline > gpurun_out/bench_p100_share$s.json 2> gpurun_out/bench_p100_share$s.err
  python - <<PY
import json
f='gpurun_out/bench_p100_share$s.json'
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print('share $s', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4))
except Exception as e:
    print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
done
