#!/bin/bash
# persistent LSTM backward kernel: correctness vs the per-step kernels and nn.LSTM autograd, timing, training step
timeout 300 python -m pytest tests/test_gpu_backward.py -q -s -k "lstm" 2>&1 | tail -30 > gpurun_out/t_lstm_bwd.log
timeout 100 python profiles/lstm_bwd_time.py > gpurun_out/lstm_bwd_time.txt 2>&1
VOG_B200_SO=$PWD/vognet_pytorch_b200/libvog_b200_trace.so timeout 100 python profiles/lstm_bwd_time.py > gpurun_out/lstm_bwd_trace.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_gpu_train_step.py tests/test_gpu_tc_bwd.py -x -q 2>&1 | tail -5 > gpurun_out/t_train_all.log
timeout 300 python bench.py --train --no-extras --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_p100_lb.json 2> gpurun_out/train_p100_lb.err
