#!/bin/bash
mkdir -p gpurun_out
export NCONF=1 ITERS=3 CHECK=0
for d in 31 95 223 479 991 1023; do echo "DBG=$d"; MB200_FA_BWD_DBG=$d timeout 100 python tools/attn_bwd_once.py 2>&1 | grep "bwd ours" | sed 's/.*bwd ours/bwd ours/' | cut -c1-120; done 2>&1 | tee gpurun_out/r2_bwd_dbg4.log
