#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_extra_kernels_gpu.py -q -m gpu -k nvfp4 > gpurun_out/nvfp4_test.log 2>&1; echo "nvfp4 tests rc=$?"; grep -E "passed|failed|Error|FAILED|max err|assert" gpurun_out/nvfp4_test.log | tail -12 | cut -c1-300
timeout 300 python tools/mxfp8_bench.py > gpurun_out/mxfp8_bench2.log 2>&1; tail -6 gpurun_out/mxfp8_bench2.log | cut -c1-400
