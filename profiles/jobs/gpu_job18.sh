#!/bin/bash
mkdir -p gpurun_out
timeout 200 python profiles/attn_ab.py > gpurun_out/attn_ab_spin.log 2>&1
cat gpurun_out/attn_ab_spin.log
