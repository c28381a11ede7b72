#!/bin/bash
# store-path experiments on the S/dP kernel (timing only; results of the variants are not valid)
for V in AB_EXP_NO_STORE AB_EXP_TILE_MAJOR; do
  VOG_NVCC_EXTRA="-D$V" python -c "from vognet_pytorch_b200 import _lib; _lib.build(force=True)"
  echo "== $V" >> gpurun_out/sdp_exp.txt
  python profiles/one_op.py attn_bwd 40 2000 768 >> gpurun_out/sdp_exp.txt 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_attn_bwd_sdp -c 2 python profiles/one_op.py attn_bwd 40 2000 768 2>&1 | grep -E "gpu__time_duration" >> gpurun_out/sdp_exp.txt
done
