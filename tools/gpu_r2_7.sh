#!/bin/bash
mkdir -p gpurun_out
( MB200_DEBUG_TMAP=1 timeout 600 python -m pytest tests/test_attn_gpu.py -q -x -s ) > gpurun_out/r2_attn_tests.log 2>&1; echo "attn tests rc=$?"; grep -E "mb200|passed|failed" gpurun_out/r2_attn_tests.log | head
