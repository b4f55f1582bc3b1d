#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"gemm_(mxfp8|nvfp4)_kernel" --launch-skip 4 --launch-count 2 -f -o gpurun_out/prof_lowp_gemm python tools/lowp_gemm_once.py > gpurun_out/ncu_lowp.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_lowp.log
ls -la gpurun_out/prof_lowp_gemm.ncu-rep
