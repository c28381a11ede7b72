#!/bin/bash
# programmatic dependent launch A/B on the headline workload + correctness of the graph / eager paths
timeout 600 python -m pytest tests/test_gpu_model_tc.py tests/test_gpu_lstm.py tests/test_gpu_tc_gemm.py tests/test_gpu_tc_attn.py -x -q 2>&1 | tail -4 > gpurun_out/t_pdl.log
for V in 1 0 1 0; do
  VOG_PDL=$V python bench.py --no-extras --no-cpu-baseline --no-seq4000 --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$V', 'ms', round(d['ms_per_step'],4), 'median', round(d['ms_per_step_median'],4), 'min', round(d['ms_per_step_min'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']))" >> gpurun_out/t_pdl.log
done
VOG_PDL=1 python bench.py --workload spat_p100 --no-extras --no-cpu-baseline --no-seq4000 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('p100 PDL=1', 'ms', round(d['ms_per_step'],4), 'value', round(d['value']))" >> gpurun_out/t_pdl.log
VOG_PDL=0 python bench.py --workload spat_p100 --no-extras --no-cpu-baseline --no-seq4000 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('p100 PDL=0', 'ms', round(d['ms_per_step'],4), 'value', round(d['value']))" >> gpurun_out/t_pdl.log
