#!/bin/bash
# per-kernel durations of ONE training step (ncu serialises and runs cold: use the SHARES, not the absolutes)
for W in spat_p100 spat_gt5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_${W}.csv \
      python bench.py --train --workload $W --steps 1 --warmup 3 > gpurun_out/ncu_train_${W}.log 2>&1
  python profiles/summarize_launches.py gpurun_out/launches_train_${W}.csv > gpurun_out/launches_train_${W}.txt 2>&1
done
