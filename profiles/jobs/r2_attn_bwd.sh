#!/bin/bash
# timing of the attention training forward / backward at the spat/p100 shapes + one ncu --set full capture of the
# S/dP kernel (mul shape)
for A in "40 2000 768" "40 2000 768 drop" "4 4000 512" "4 4000 512 drop"; do python profiles/one_op.py attn_bwd $A; done > gpurun_out/attn_bwd_times.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_attn_bwd_sdp -c 1 -f -o gpurun_out/attn_bwd_sdp_mul python profiles/one_op.py attn_bwd 40 2000 768 drop > gpurun_out/ncu_attn_bwd.log 2>&1
