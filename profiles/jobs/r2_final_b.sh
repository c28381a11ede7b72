#!/bin/bash
# full GPU suite + the driver's bench line + spat/p100 forward and training-step lines (after the record exchange and
# PredictionFetcher changes)
rm -f gpurun_out/parity_margins.json
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t_all_r2b.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err
timeout 300 python bench.py --workload spat_p100 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_p100_c.json 2> gpurun_out/bench_p100_c.err
