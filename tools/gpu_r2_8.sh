#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2_gpu_suite.log
timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n1.json | cut -c1-900; tail -3 gpurun_out/r2_bench_n1.err | cut -c1-400
MEGATRON_B200_ATTN_BWD=library timeout 400 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_bench_n1_libbwd.json 2> gpurun_out/r2_bench_n1_libbwd.err; echo "bench(lib bwd) rc=$?"; tail -1 gpurun_out/r2_bench_n1_libbwd.json | cut -c1-400
