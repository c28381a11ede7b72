#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -x -q > gpurun_out/pytest_lstm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lstm.log
timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_notrace.log 2>&1
tail -3 gpurun_out/pytest_lstm.log; grep median gpurun_out/lstm_notrace.log
