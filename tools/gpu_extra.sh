#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_extra_kernels_gpu.py -q -m gpu -x --deselect tests/test_extra_kernels_gpu.py::test_mxfp8_block_scaled_gemm > gpurun_out/extra_test.log 2>&1; echo "extra kernels rc=$?"; tail -25 gpurun_out/extra_test.log | cut -c1-400
timeout 300 python -m pytest tests/test_extra_kernels_gpu.py -q -m gpu -k mxfp8_block_scaled > gpurun_out/mxfp8_gemm_test.log 2>&1; echo "mxfp8 gemm rc=$?"; tail -25 gpurun_out/mxfp8_gemm_test.log | cut -c1-400
