#!/bin/bash
# record exchange of the recurrence kernel (xmode 2): correctness, phase trace, 144-register variant, effect on the gt5 step
timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q 2>&1 | tail -5 > gpurun_out/t_lstm_rec.log
timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_rec_time.txt 2>&1
VOG_B200_SO=$PWD/vognet_pytorch_b200/libvog_b200_trace.so timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_rec_trace.txt 2>&1
VOG_B200_SO=$PWD/vognet_pytorch_b200/libvog_b200_nreg3.so timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_rec_time_nreg3.txt 2>&1
VOG_B200_SO=$PWD/vognet_pytorch_b200/libvog_b200_nreg3trace.so timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_rec_trace_nreg3.txt 2>&1
for X in 0 2; do
  VOG_LSTM_XMODE=$X timeout 200 python bench.py --no-extras --no-cpu-baseline --no-seq4000 --steps 100 --warmup 10 > gpurun_out/bench_gt5_x$X.json 2> gpurun_out/bench_gt5_x$X.err
done
VOG_LSTM_XMODE=2 VOG_B200_SO=$PWD/vognet_pytorch_b200/libvog_b200_nreg3.so timeout 200 python bench.py --no-extras --no-cpu-baseline --no-seq4000 --steps 100 --warmup 10 > gpurun_out/bench_gt5_x2_nreg3.json 2> gpurun_out/bench_gt5_x2_nreg3.err
