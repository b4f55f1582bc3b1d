#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_attn_gpu.py -q -x ) > gpurun_out/r2_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -3 gpurun_out/r2_attn_tests.log | cut -c1-300
timeout 300 python tools/attn_bwd_once.py 2>&1 | tee gpurun_out/r2_attn_bwd.log | cut -c1-300
export NCONF=1 ITERS=3 CHECK=0
for d in 31 1 5 12; do echo "DBG=$d"; MB200_FA_BWD_DBG=$d timeout 100 python tools/attn_bwd_once.py 2>&1 | grep heads | sed 's/.*bwd ours/bwd ours/' | cut -c1-120; done 2>&1 | tee gpurun_out/r2_bwd_dbg3.log
