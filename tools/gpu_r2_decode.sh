#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_paged_attention_gpu.py -q -x ) > gpurun_out/r2_paged_tests.log 2>&1; echo "paged tests rc=$?"; tail -4 gpurun_out/r2_paged_tests.log | cut -c1-400
timeout 400 python tools/decode_bench.py > gpurun_out/r2_decode_bench.log 2>&1; echo "decode bench rc=$?"; grep bench gpurun_out/r2_decode_bench.log | cut -c1-400; tail -3 gpurun_out/r2_decode_bench.log | cut -c1-300
