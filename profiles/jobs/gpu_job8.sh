#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_attn.py -x -q > gpurun_out/pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attn.log
timeout 200 python profiles/attn_ab.py > gpurun_out/attn_ab.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_attn.log; cat gpurun_out/attn_ab.log; tail -3 gpurun_out/pytest_gpu.log
