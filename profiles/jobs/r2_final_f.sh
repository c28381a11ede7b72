#!/bin/bash
# projection + key-factor expansion in one launch (two launches fewer per forward): full GPU suite + the driver's bench line
rm -f gpurun_out/parity_margins.json
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t_all_r2f.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_g.json 2> gpurun_out/bench_r2_g.err
