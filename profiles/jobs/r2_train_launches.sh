#!/bin/bash
# per-kernel durations of ONE training step (ncu serialises and runs cold: use the SHARES, not the absolutes)
for W in spat_p100; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_${W}.csv \
      python bench.py --train --no-extras --workload $W --steps 1 --warmup 3 > gpurun_out/ncu_train_${W}.log 2>&1
  python profiles/summarize_launches.py gpurun_out/launches_train_${W}.csv > gpurun_out/launches_train_${W}.txt 2>&1
done
timeout 300 python bench.py --train --no-extras --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_p100_n1.json 2> gpurun_out/train_p100_n1.err
timeout 200 python -m pytest tests/test_gpu_backward.py -x -q -k "dropout" 2>&1 | tail -3 > gpurun_out/t_drop.log
