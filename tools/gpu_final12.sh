#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_bwd --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_fa_bwd python tools/attn_bwd_once.py > gpurun_out/ncu_fa_bwd.log 2>&1; echo "ncu fa bwd rc=$?"; tail -2 gpurun_out/ncu_fa_bwd.log
