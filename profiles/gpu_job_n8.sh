#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N" > gpurun_out/n8_info.txt
for n in 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/bench_n${n}_spat_gt5.json 2> gpurun_out/bench_n${n}_spat_gt5.err
    tail -1 gpurun_out/bench_n${n}_spat_gt5.json | cut -c1-200
  fi
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 --workload spat_p100 > gpurun_out/bench_n8_spat_p100.json 2> gpurun_out/bench_n8_spat_p100.err
tail -1 gpurun_out/bench_n8_spat_p100.json | cut -c1-200
tail -2 gpurun_out/bench_n8_spat_gt5.err
