#!/bin/bash
# 2-GPU: fused TP GEMM kernels (numerics + timing), collective microbench, GEMM epilogue retest, N=2 benches
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/fused_tp_test.py > gpurun_out/fused_tp_test.log 2>&1; echo "fused test rc=$?"; tail -60 gpurun_out/fused_tp_test.log
timeout 240 $TR --master-port 29512 tools/nvl_coll_bench.py > gpurun_out/nvl_coll_bench.log 2>&1; echo "coll bench rc=$?"; tail -40 gpurun_out/nvl_coll_bench.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu > gpurun_out/gemm_test2.log 2>&1; echo "gemm test rc=$?"; tail -5 gpurun_out/gemm_test2.log
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench2.log 2>&1; echo "gemm bench rc=$?"; tail -40 gpurun_out/gemm_bench2.log
for mode in fused nccl; do
  timeout 600 $TR --master-port 29533 bench.py --gpus ${NG:-2} --steps 3 --warmup 3 --no-e2e --tp-comm $mode > gpurun_out/bench_n${NG:-2}_$mode.log 2>&1
  echo "bench $mode rc=$?"; grep '^{' gpurun_out/bench_n${NG:-2}_$mode.log || tail -25 gpurun_out/bench_n${NG:-2}_$mode.log
done
