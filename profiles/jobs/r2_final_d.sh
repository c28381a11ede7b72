#!/bin/bash
# final round-2 evidence: full GPU suite, the driver's bench line, launch lists (forward gt5 + training step p100, eager:
# graph nodes with tcgen05 kernels cannot be profiled), ncu --set full of the two-units-per-warp recurrence kernel,
# spat/p100 forward and training-step lines
rm -f gpurun_out/parity_margins.json
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t_all_r2d.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_spat_gt5.csv \
    python bench.py --no-graph --no-extras --no-cpu-baseline --no-seq4000 --steps 2 --warmup 3 > gpurun_out/ncu_r2_gt5.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_r2_spat_gt5.csv > gpurun_out/launches_r2_spat_gt5.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_pair -s 2 -c 1 -f -o gpurun_out/lstm_pair python profiles/one_op.py lstm > gpurun_out/ncu_lstm_pair.log 2>&1
python profiles/ncu_summary.py gpurun_out/lstm_pair.ncu-rep > gpurun_out/ncu_lstm_pair.txt 2>&1
timeout 300 python bench.py --workload spat_p100 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_p100_e.json 2> gpurun_out/bench_p100_e.err
timeout 300 python bench.py --train --no-extras --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_p100_e.json 2> gpurun_out/train_p100_e.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train_spat_p100.csv \
    python bench.py --train --no-extras --workload spat_p100 --steps 1 --warmup 2 > gpurun_out/ncu_train_spat_p100.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_train_spat_p100.csv > gpurun_out/launches_train_spat_p100.txt 2>&1
