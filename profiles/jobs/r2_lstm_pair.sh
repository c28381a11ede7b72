#!/bin/bash
# two hidden units per warp in the recurrence kernel (xmode 3): correctness, phase trace, effect on the gt5 step
timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q 2>&1 | tail -5 > gpurun_out/t_lstm_pair.log
timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_pair_time.txt 2>&1
VOG_B200_SO=$PWD/vognet_pytorch_b200/libvog_b200_trace.so timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_pair_trace.txt 2>&1
for X in 2 3; do
  VOG_LSTM_XMODE=$X timeout 200 python bench.py --no-extras --no-cpu-baseline --no-seq4000 --steps 100 --warmup 10 > gpurun_out/bench_gt5_x$X.json 2> gpurun_out/bench_gt5_x$X.err
done
