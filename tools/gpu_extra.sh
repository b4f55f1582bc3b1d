#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_extra_kernels_gpu.py -q -m gpu -k mxfp8 > gpurun_out/extra_test2.log 2>&1; echo "mxfp8 tests rc=$?"; grep -E "passed|failed|Error|FAILED" gpurun_out/extra_test2.log | tail -25 | cut -c1-300
timeout 300 python tools/mxfp8_bench.py > gpurun_out/mxfp8_bench.log 2>&1; cat gpurun_out/mxfp8_bench.log | tail -8
