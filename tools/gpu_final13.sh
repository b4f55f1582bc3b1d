#!/bin/bash
# final sanity: whole single-GPU test suite, default bench, bench with the residual-fused RMSNorm layer path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --ignore=tests/test_nvlink_gpu.py --ignore=tests/test_nvlink_moe_gpu.py --ignore=tests/test_moe_e2e_gpu.py > gpurun_out/final_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/final_gpu_suite.log | cut -c1-300
timeout 400 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err; echo "bench default rc=$?"; tail -1 gpurun_out/bench_final_default.json | cut -c1-700
MEGATRON_B200_FUSED_RESIDUAL_NORM=1 timeout 400 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_final_fusednorm.json 2> gpurun_out/bench_final_fusednorm.err; echo "bench fused-norm rc=$?"; tail -1 gpurun_out/bench_final_fusednorm.json | cut -c1-400
