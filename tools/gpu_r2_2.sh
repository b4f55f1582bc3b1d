#!/bin/bash
# round 2, call 2 (2 GPUs): multi-GPU tests (NVLS multimem kernels, fused pair ops incl. GEMM->AR), bench both arms at N=2
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_nvlink_gpu.py tests/test_nvlink_moe_gpu.py tests/test_moe_e2e_gpu.py -q -x -s ) > gpurun_out/r2_mg_tests_n2.log 2>&1; echo "mg tests rc=$?"; tail -6 gpurun_out/r2_mg_tests_n2.log | cut -c1-1500
P=29600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"; tail -1 gpurun_out/r2_bench_n2.json | cut -c1-3000; tail -5 gpurun_out/r2_bench_n2.err | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref_n2.json | cut -c1-2000; tail -5 gpurun_out/r2_bench_ref_n2.err | cut -c1-600
