#!/bin/bash
# 8 GPUs: BASELINE config 5 - spat/p100 training step, bs=32 global, NCCL all-reduce of the flat gradient
nvidia-smi -L > gpurun_out/n8_gpus.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 8 --train --workload spat_p100 --steps 10 --warmup 3 > gpurun_out/train_spat_p100_n8.json 2> gpurun_out/train_spat_p100_n8.err
echo "rc=$?" >> gpurun_out/train_spat_p100_n8.err
