#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemm.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-seq4000 > gpurun_out/bench_spat_gt5.json 2> gpurun_out/bench_spat_gt5.err
timeout 300 python bench.py --workload spat_p100 --steps 10 --no-seq4000 > gpurun_out/bench_spat_p100.json 2> gpurun_out/bench_spat_p100.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_spat_gt5.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_gt5.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_spat_p100.csv python bench.py --workload spat_p100 --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-seq4000 > gpurun_out/ncu_p100.log 2>&1
VOG_NVCC_EXTRA=-DVOG_ATTN_PROFILE python -c "from vognet_pytorch_b200 import _lib; _lib.build(force=True)" > gpurun_out/rebuild.log 2>&1
timeout 200 python profiles/attn_phases.py > gpurun_out/attn_phases.log 2>&1
tail -3 gpurun_out/pytest_gemm.log; tail -3 gpurun_out/pytest_gpu.log; cut -c1-330 gpurun_out/bench_spat_gt5.json; cut -c1-330 gpurun_out/bench_spat_p100.json; cat gpurun_out/attn_phases.log
