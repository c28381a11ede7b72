#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -x -q > gpurun_out/pytest_lstm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lstm.log
timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_notrace.log 2>&1
VOG_NVCC_EXTRA=-DVOG_LSTM_TRACE python -c "from vognet_pytorch_b200 import _lib; _lib.build(force=True)" > gpurun_out/rebuild.log 2>&1
timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_trace.log 2>&1
tail -3 gpurun_out/pytest_lstm.log; grep median gpurun_out/lstm_notrace.log; grep "per step" gpurun_out/lstm_trace.log
