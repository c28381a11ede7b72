#!/bin/bash
# 2 GPUs: NCCL tests (gradient all-reduce, prediction gather), the driver's bench line under torchrun, training step
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -k nccl 2>&1 | tail -8 > gpurun_out/t_n2_trainstep.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err
echo "rc=$?" >> gpurun_out/bench_n2_r2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
   bench.py --gpus 2 --train --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_spat_p100_n2.json 2> gpurun_out/train_spat_p100_n2.err
echo "rc=$?" >> gpurun_out/train_spat_p100_n2.err
