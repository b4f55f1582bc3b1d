#!/bin/bash
mkdir -p gpurun_out
export NCONF=1 ITERS=2 CHECK=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_bwd_dkv -s 2 -c 1 -o gpurun_out/r2_fa_bwd_dkv2 python tools/attn_bwd_once.py > gpurun_out/r2_ncu_dkv2.log 2>&1; echo "ncu rc=$?"
