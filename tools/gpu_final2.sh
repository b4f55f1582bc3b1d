#!/bin/bash
# 1-GPU: rest of the GPU suite (no -x), attention probe with the two-issuer kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/final_gpu_tests2.log 2>&1; echo "gpu tests rc=$?"; tail -8 gpurun_out/final_gpu_tests2.log | cut -c1-300
timeout 300 python tools/attn_probe.py 2>&1 | grep "megatron_b200\|sdpa-cudnn" > gpurun_out/attn_probe5.log; cat gpurun_out/attn_probe5.log
