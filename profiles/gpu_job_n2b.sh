#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2_spat_gt5.json 2> gpurun_out/bench_n2_spat_gt5.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --workload spat_p100 --no-seq4000 > gpurun_out/bench_n2_spat_p100.json 2> gpurun_out/bench_n2_spat_p100.err
for f in gpurun_out/bench_n2_spat_gt5.json gpurun_out/bench_n2_spat_p100.json; do echo $f; tail -1 $f | cut -c1-200; done
tail -3 gpurun_out/bench_n2_spat_gt5.err
