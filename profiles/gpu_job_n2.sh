#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2_spat_gt5.json 2> gpurun_out/bench_n2_spat_gt5.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --workload temp_gt5 > gpurun_out/bench_n2_temp_gt5.json 2> gpurun_out/bench_n2_temp_gt5.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload spat_p100 > gpurun_out/bench_n2_spat_p100.json 2> gpurun_out/bench_n2_spat_p100.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
for f in gpurun_out/bench_n2_*.json; do echo $f; tail -1 $f | cut -c1-260; done
tail -3 gpurun_out/bench_n2_spat_gt5.err
