#!/bin/bash
# Round-end evidence refresh (trimmed): parity tests, bench lines, reference arm, launch lists, ncu capture of the
# device-side concatenation kernel.  The ncu --set full captures of attention / GEMM / LSTM are unchanged since
# profiles/gpu_job_final.sh last ran (no kernel on the BASELINE path changed).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 300 python bench.py --workload temp_gt5 --no-seq4000 > gpurun_out/bench_temp_gt5.json 2> gpurun_out/bench_temp_gt5.err
timeout 300 python bench.py --workload spat_p100 --steps 30 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
timeout 200 python bench.py --impl reference --steps 5 > gpurun_out/bench_ref_gt5.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_spat_gt5.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_gt5.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_spat_p100.csv python bench.py --workload spat_p100 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_p100.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:permute_rows -s 3 -c 1 -o gpurun_out/relayout_feat_p100 python profiles/relayout_bw.py > gpurun_out/ncu_relayout.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ('gpurun_out/bench_spat_gt5.json','gpurun_out/bench_temp_gt5.json','gpurun_out/bench_spat_p100.json','gpurun_out/bench_ref_gt5.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4), 'launches', d.get('gpu_launches'), 'roof', (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_seq4000') or {}).get('frac'), 'clocks', d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
