#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_graph_gpu.py tests/test_attn_gpu.py -q -m gpu > gpurun_out/final_gpu_tests3.log 2>&1; echo "graph+attn tests rc=$?"; tail -4 gpurun_out/final_gpu_tests3.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_fa_v4 python tools/attn_once.py > gpurun_out/ncu_fa_v4.log 2>&1; echo "ncu fa rc=$?"; tail -2 gpurun_out/ncu_fa_v4.log
