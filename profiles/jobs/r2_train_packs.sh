#!/bin/bash
# single-pass weight re-packing + direct accumulation of the LSTM weight gradients: training tests, training-step lines
timeout 400 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train_step.py tests/test_gpu_tc_bwd.py tests/test_optim.py -x -q 2>&1 | tail -5 > gpurun_out/t_train_packs.log
timeout 300 python bench.py --train --no-extras --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_p100_pk.json 2> gpurun_out/train_p100_pk.err
timeout 300 python bench.py --train --no-extras --workload spat_gt5 --steps 10 --warmup 3 > gpurun_out/train_gt5_pk.json 2> gpurun_out/train_gt5_pk.err
