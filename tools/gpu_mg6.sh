#!/bin/bash
# N-GPU (NG=4 or 8): bench in nccl mode (+fused when FUSED=1), pair-op timings, NVLink collective + MoE tests
mkdir -p gpurun_out
N=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --tp-comm nccl > gpurun_out/bench_n${N}_nccl.log 2>&1; echo "bench nccl rc=$?"; grep '^{' gpurun_out/bench_n${N}_nccl.log | cut -c1-900 || tail -25 gpurun_out/bench_n${N}_nccl.log
MODES=${MODES:-nccl,fused:4:8,fused:8:12} ITERS=10 timeout 300 $TR --master-port 29511 tools/fused_tp_test.py > gpurun_out/fused_tp_test_n$N.log 2>&1; echo "fused test rc=$?"; grep -v "^\*\|OMP" gpurun_out/fused_tp_test_n$N.log | tail -60
if [ "${FUSED:-1}" = "1" ]; then
  timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --tp-comm fused > gpurun_out/bench_n${N}_fused.log 2>&1; echo "bench fused rc=$?"; grep '^{' gpurun_out/bench_n${N}_fused.log | cut -c1-900 || tail -25 gpurun_out/bench_n${N}_fused.log
fi
timeout 300 python -m pytest tests/test_nvlink_gpu.py tests/test_nvlink_moe_gpu.py -q -m gpu > gpurun_out/nvlink_tests_n$N.log 2>&1; echo "nvlink tests rc=$?"; tail -15 gpurun_out/nvlink_tests_n$N.log | cut -c1-300
