#!/bin/bash
# 8 GPUs: the driver's bench line (forward, weak scaling) with its train_step sub-object (config 5: spat/p100 training,
# bs=32 over 8 GPUs, NCCL gradient all-reduce) + the training bench alone
nvidia-smi -L > gpurun_out/n8_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_r2.json 2> gpurun_out/bench_n8_r2.err
echo "rc=$?" >> gpurun_out/bench_n8_r2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
   bench.py --gpus 8 --train --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_spat_p100_n8.json 2> gpurun_out/train_spat_p100_n8.err
echo "rc=$?" >> gpurun_out/train_spat_p100_n8.err
