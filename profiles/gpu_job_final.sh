#!/bin/bash
# Round-end evidence refresh: parity tests, bench lines, launch lists, ncu captures of the dominant kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 300 python bench.py --workload temp_gt5 --no-seq4000 > gpurun_out/bench_temp_gt5.json 2> gpurun_out/bench_temp_gt5.err
timeout 300 python bench.py --workload spat_p100 --steps 30 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
timeout 200 python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_gt5.json 2>&1
timeout 200 python profiles/microbench.py p100 > gpurun_out/micro_p100.log 2>&1
timeout 200 python profiles/microbench.py gt5 > gpurun_out/micro_gt5.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_spat_gt5.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_gt5.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_spat_p100.csv python bench.py --workload spat_p100 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_p100.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_attn2 -s 2 -c 1 -o gpurun_out/attn2_mul_p100 python profiles/one_op.py attn 40 2000 768 > gpurun_out/ncu_attn_mul.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_attn2 -s 2 -c 1 -o gpurun_out/attn2_obj_p100 python profiles/one_op.py attn 4 4000 512 > gpurun_out/ncu_attn_obj.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/gemm_qkvf_p100 python profiles/one_op.py qkvf 4 10 5 400 > gpurun_out/ncu_qkvf.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/gemm_gres_p100 python profiles/one_op.py gres 4 10 5 400 > gpurun_out/ncu_gres.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ('gpurun_out/bench_spat_gt5.json','gpurun_out/bench_temp_gt5.json','gpurun_out/bench_spat_p100.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d['gpu_launches'], 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4), (d.get('roofline_seq4000') or {}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-1500:])
PY
