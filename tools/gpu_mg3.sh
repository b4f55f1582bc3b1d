#!/bin/bash
# 2-GPU: fused TP GEMM (unrolled comm) sweep + collective bench; 1-GPU: flash attention + grouped GEMM + cuda graph tests, attention probe
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_cuda_graph_gpu.py tests/test_moe_gpu.py -x -q -m gpu > gpurun_out/attn_test.log 2>&1; echo "attn/graph/moe test rc=$?"; tail -30 gpurun_out/attn_test.log
timeout 300 python tools/attn_probe.py > gpurun_out/attn_probe2.log 2>&1; echo "attn probe rc=$?"; tail -12 gpurun_out/attn_probe2.log
timeout 400 $TR --master-port 29511 tools/fused_tp_test.py > gpurun_out/fused_tp_test2.log 2>&1; echo "fused test rc=$?"; tail -75 gpurun_out/fused_tp_test2.log
timeout 240 $TR --master-port 29512 tools/nvl_coll_bench.py > gpurun_out/nvl_coll_bench2.log 2>&1; echo "coll bench rc=$?"; tail -40 gpurun_out/nvl_coll_bench2.log
