#!/bin/bash
# one ncu --set full capture of the S/dP kernel (mul shape, dropout on / off)
ncu --set full --clock-control none --import-source on -k regex:tc_attn_bwd_sdp -c 1 -f -o gpurun_out/attn_bwd_sdp_mul python profiles/one_op.py attn_bwd 40 2000 768 > gpurun_out/ncu_attn_bwd.log 2>&1
