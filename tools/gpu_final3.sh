#!/bin/bash
# N-GPU: exposed communication per step (pair ops vs GEMM-only floor) and the default (auto) bench
mkdir -p gpurun_out
N=${NG:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
MODES=nccl,local,fused:4:8 ITERS=10 timeout 300 $TR --master-port 29511 tools/fused_tp_test.py > gpurun_out/exposed_comm_n$N.log 2>&1; echo "exposed comm rc=$?"; grep -v "^\*\|OMP" gpurun_out/exposed_comm_n$N.log | grep "exposed\|mode" | cut -c1-200
if [ "${BENCH:-1}" = "1" ]; then
  timeout 600 $TR --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_auto.log 2>&1; echo "bench auto rc=$?"; grep '^{' gpurun_out/bench_n${N}_auto.log | cut -c1-300 || tail -25 gpurun_out/bench_n${N}_auto.log
fi
