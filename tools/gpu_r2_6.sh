#!/bin/bash
mkdir -p gpurun_out
export NCONF=1 ITERS=3 CHECK=0
for d in 0 1 2 4 8 3 12 5 15; do echo "DBG=$d"; MB200_FA_BWD_DBG=$d timeout 100 python tools/attn_bwd_once.py 2>&1 | grep heads | sed 's/.*bwd ours/bwd ours/' | cut -c1-120; done 2>&1 | tee gpurun_out/r2_bwd_dbg.log
