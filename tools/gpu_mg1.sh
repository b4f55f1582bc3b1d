#!/bin/bash
# 2-GPU: NVLink collectives vs NCCL, then short N=2 benches (nccl vs nvlink)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 420 python -m pytest tests/test_nvlink_gpu.py -x -q -m gpu -s > gpurun_out/nvlink_test.log 2>&1; echo "nvlink test rc=$?"; tail -25 gpurun_out/nvlink_test.log
for mode in nccl nvlink; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --tp-comm $mode > gpurun_out/bench_n2_$mode.log 2>&1
  echo "bench n2 $mode rc=$?"; grep '^{' gpurun_out/bench_n2_$mode.log || tail -25 gpurun_out/bench_n2_$mode.log
done
