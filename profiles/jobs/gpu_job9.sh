#!/bin/bash
mkdir -p gpurun_out
VOG_NVCC_EXTRA=-DVOG_ATTN_PROFILE python -c "from vognet_pytorch_b200 import _lib; _lib.build(force=True)" > gpurun_out/rebuild.log 2>&1
timeout 200 python profiles/attn_phases.py > gpurun_out/attn_phases_v2.log 2>&1
cat gpurun_out/attn_phases_v2.log
