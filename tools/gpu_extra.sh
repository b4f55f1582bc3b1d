#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --ignore=tests/test_nvlink_gpu.py --ignore=tests/test_nvlink_moe_gpu.py --ignore=tests/test_moe_e2e_gpu.py > gpurun_out/final_gpu_suite2.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/final_gpu_suite2.log | tail -12 | cut -c1-300
