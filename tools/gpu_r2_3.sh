#!/bin/bash
# round 2, call 3 (1 GPU): new two-kernel attention backward — correctness, then timing at TP=1/2/4/8 head counts
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_attn_gpu.py -q -x ) > gpurun_out/r2_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -15 gpurun_out/r2_attn_tests.log | cut -c1-400
timeout 300 python tools/attn_bwd_once.py > gpurun_out/r2_attn_bwd.log 2>&1; echo "bwd timing rc=$?"; tail -12 gpurun_out/r2_attn_bwd.log | cut -c1-400
