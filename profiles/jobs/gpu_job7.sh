#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/gemm_qkvf_p100 python profiles/one_op.py qkvf 4 10 5 400 > gpurun_out/ncu_qkvf.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 1 -o gpurun_out/gemm_gres_p100 python profiles/one_op.py gres 4 10 5 400 > gpurun_out/ncu_gres.log 2>&1
tail -2 gpurun_out/ncu_qkvf.log gpurun_out/ncu_gres.log
