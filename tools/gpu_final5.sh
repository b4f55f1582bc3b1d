#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_attn_gpu.py -q -m gpu > gpurun_out/attn_test_ts.log 2>&1; echo "attn tests (both variants) rc=$?"; tail -12 gpurun_out/attn_test_ts.log | cut -c1-300
timeout 300 python tools/attn_probe.py 2>&1 | grep "megatron_b200\|sdpa-cudnn" > gpurun_out/attn_probe6.log; cat gpurun_out/attn_probe6.log
