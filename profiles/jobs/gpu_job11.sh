#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 300 python bench.py --workload spat_p100 --steps 10 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_resident -s 2 -c 1 -o gpurun_out/lstm_resident python profiles/one_op.py lstm > gpurun_out/ncu_lstm.log 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/bench_spat_gt5.json','gpurun_out/bench_spat_p100.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],4), 'attn', round(d['roofline_attention']['frac'],4), d.get('roofline_seq4000',{}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-800:])
PY
