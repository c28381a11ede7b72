#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py -x -q > gpurun_out/pytest_lstm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lstm.log
timeout 200 python profiles/gemm_trace.py > gpurun_out/gemm_trace.log 2>&1
timeout 100 python profiles/lstm_trace.py > gpurun_out/lstm_trace.log 2>&1
timeout 200 python profiles/microbench.py p100 > gpurun_out/micro_p100.log 2>&1
timeout 300 python bench.py --workload spat_p100 --steps 10 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
timeout 200 python profiles/microbench.py gt5 > gpurun_out/micro_gt5.log 2>&1
timeout 300 python bench.py --no-seq4000 > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_lstm.log; cat gpurun_out/gemm_trace.log gpurun_out/lstm_trace.log gpurun_out/micro_gt5.log gpurun_out/micro_p100.log; cut -c1-330 gpurun_out/bench_spat_gt5.json; cut -c1-330 gpurun_out/bench_spat_p100.json;  tail -3 gpurun_out/pytest_gpu.log
