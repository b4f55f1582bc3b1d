#!/bin/bash
mkdir -p gpurun_out
export NCONF=1 ITERS=2 CHECK=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_ --csv --log-file gpurun_out/r2_fa_launches.csv python tools/attn_bwd_once.py > gpurun_out/r2_fa_launches.log 2>&1; echo "rc=$?"
grep -E "fa_" gpurun_out/r2_fa_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_bwd_dkv -s 2 -c 1 -o gpurun_out/r2_fa_bwd_dkv python tools/attn_bwd_once.py > gpurun_out/r2_ncu_dkv.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_bwd_dq -s 2 -c 1 -o gpurun_out/r2_fa_bwd_dq python tools/attn_bwd_once.py > gpurun_out/r2_ncu_dq.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
