#!/bin/bash
# 2-GPU box: N=2 bench fused vs nccl with the current kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in fused nccl; do
  timeout 600 $TR --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --tp-comm $mode > gpurun_out/bench_n2_${mode}_r2.log 2>&1; echo "bench n2 $mode rc=$?"; grep '^{' gpurun_out/bench_n2_${mode}_r2.log | cut -c1-420 || tail -20 gpurun_out/bench_n2_${mode}_r2.log
done
timeout 600 python -m pytest tests/test_moe_e2e_gpu.py -q -m gpu > gpurun_out/moe_e2e.log 2>&1; echo "moe e2e rc=$?"; tail -25 gpurun_out/moe_e2e.log | cut -c1-300
