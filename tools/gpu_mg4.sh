#!/bin/bash
# 2-GPU: fused TP GEMM (warp-granular comm) + reference arm at N=2; 1-GPU: cuda graph / moe / fp8 tests, attention probe
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_cuda_graph_gpu.py tests/test_moe_gpu.py tests/test_attn_gpu.py "tests/test_gemm_gpu.py::test_fp8_gemm_matches_emulation" -q -m gpu > gpurun_out/misc_test.log 2>&1; echo "graph/moe/attn/fp8 test rc=$?"; tail -40 gpurun_out/misc_test.log
timeout 300 python tools/attn_probe.py 2>&1 | grep -v Warn | grep "TF\|FAILED" > gpurun_out/attn_probe3.log; cat gpurun_out/attn_probe3.log
MODES=nccl,fused:4:8,fused:8:12 timeout 400 $TR --master-port 29511 tools/fused_tp_test.py > gpurun_out/fused_tp_test3.log 2>&1; echo "fused test rc=$?"; grep -v "^\*\|OMP" gpurun_out/fused_tp_test3.log | tail -60
timeout 900 $TR --master-port 29533 bench.py --impl reference --gpus ${NG:-2} --steps 3 --warmup 3 > gpurun_out/bench_ref_n${NG:-2}.log 2>&1; echo "ref bench rc=$?"; grep '^{' gpurun_out/bench_ref_n${NG:-2}.log || tail -30 gpurun_out/bench_ref_n${NG:-2}.log
