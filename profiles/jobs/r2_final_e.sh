#!/bin/bash
# dependent launches on by default for the small configurations: full GPU suite + the driver's bench line + spat/p100
rm -f gpurun_out/parity_margins.json
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t_all_r2e.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err
timeout 300 python bench.py --workload spat_p100 --no-extras --no-cpu-baseline --no-seq4000 --steps 20 --warmup 5 > gpurun_out/bench_p100_f.json 2> gpurun_out/bench_p100_f.err
