#!/bin/bash
# the driver's smoke() entry + programmatic-dependent-launch A/B on the final kernels
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_r2.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_r2.log
for P in 0 1; do
  VOG_PDL=$P timeout 200 python bench.py --no-extras --no-cpu-baseline --no-seq4000 --steps 100 --warmup 10 > gpurun_out/bench_gt5_pdl$P.json 2> gpurun_out/bench_gt5_pdl$P.err
done
