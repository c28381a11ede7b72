#!/bin/bash
# 2 GPUs: NCCL test of the gradient all-reduce + prediction gather, then the training-step bench under torchrun
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 600 python -m pytest tests/test_gpu_train_step.py -x -q -k nccl 2>&1 | tail -60 > gpurun_out/t_n2_trainstep.log
for W in spat_p100 spat_gt5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --train --workload $W --steps 10 --warmup 3 > gpurun_out/train_${W}_n2.json 2> gpurun_out/train_${W}_n2.err
  echo "rc=$?" >> gpurun_out/train_${W}_n2.err
done
