#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1"
timeout 600 python tools/graph_probe.py > gpurun_out/graph_probe.log 2>&1; echo "graph probe rc=$?"; cat gpurun_out/graph_probe.log | cut -c1-900
timeout 900 python -m pytest tests/test_moe_gpu.py tests/test_attn_gpu.py tests/test_nvlink_moe_gpu.py "tests/test_gemm_gpu.py::test_fp8_gemm_matches_emulation" -q -m gpu > gpurun_out/misc_test2.log 2>&1; echo "moe/attn/fp8/nvlink-moe test rc=$?"; tail -40 gpurun_out/misc_test2.log | cut -c1-300
MODES=nccl,fused:4:8,fused:8:12 timeout 400 $TR --master-port 29511 tools/fused_tp_test.py > gpurun_out/fused_tp_test4.log 2>&1; echo "fused test rc=$?"; grep -v "^\*\|OMP" gpurun_out/fused_tp_test4.log | tail -56
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd --launch-skip 2 --launch-count 1 -f -o gpurun_out/prof_fa python tools/attn_once.py > gpurun_out/ncu_fa.log 2>&1; echo "ncu fa rc=$?"; tail -3 gpurun_out/ncu_fa.log
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_n1_r3.log 2>&1; echo "bench n1 rc=$?"; grep '^{' gpurun_out/bench_n1_r3.log | cut -c1-1500 || tail -30 gpurun_out/bench_n1_r3.log
