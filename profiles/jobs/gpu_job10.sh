#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 300 python bench.py --workload spat_p100 --steps 10 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
timeout 200 python profiles/microbench.py p100 > gpurun_out/micro_p100.log 2>&1
timeout 200 python profiles/microbench.py gt5 > gpurun_out/micro_gt5.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_spat_gt5.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_gt5.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_spat_p100.csv python bench.py --workload spat_p100 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_p100.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_attn2 -s 2 -c 1 -o gpurun_out/attn2_mul_p100 python profiles/one_op.py attn 40 2000 768 > gpurun_out/ncu_attn_mul.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_attn2 -s 2 -c 1 -o gpurun_out/attn2_obj_p100 python profiles/one_op.py attn 4 4000 512 > gpurun_out/ncu_attn_obj.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cut -c1-330 gpurun_out/bench_spat_gt5.json; cut -c1-330 gpurun_out/bench_spat_p100.json
python profiles/last_forward.py gpurun_out/launches_spat_gt5.csv; python profiles/last_forward.py gpurun_out/launches_spat_p100.csv
