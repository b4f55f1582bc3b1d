#!/bin/bash
# 1-GPU box: graph probe, N=1 bench with both attention implementations, step profile, attention probe, ncu of the memory-bound kernels
mkdir -p gpurun_out
timeout 900 python tools/graph_probe.py > gpurun_out/graph_probe2.log 2>&1; echo "graph probe rc=$?"; cut -c1-600 gpurun_out/graph_probe2.log
timeout 300 python tools/attn_probe.py 2>&1 | grep "megatron_b200\|sdpa-cudnn" > gpurun_out/attn_probe4.log; cat gpurun_out/attn_probe4.log
timeout 300 python -m pytest tests/test_attn_gpu.py -q -m gpu > gpurun_out/attn_test3.log 2>&1; echo "attn test rc=$?"; tail -3 gpurun_out/attn_test3.log
for impl in library native; do
  MEGATRON_B200_ATTN=$impl timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n1_attn_$impl.log 2>&1; echo "bench n1 attn=$impl rc=$?"; grep '^{' gpurun_out/bench_n1_attn_$impl.log | cut -c1-420 || tail -20 gpurun_out/bench_n1_attn_$impl.log
done
timeout 600 python tools/step_profile.py --layers 4 > gpurun_out/step_profile_n1.log 2>&1; echo "step profile rc=$?"; tail -32 gpurun_out/step_profile_n1.log | cut -c1-170
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'norm_fwd_kernel|norm_bwd_kernel|swiglu_fwd|swiglu_bwd|rope_kernel|ce_stats|ce_bwd|adam_kernel|l2norm_kernel' --launch-skip 10 --launch-count 11 -f -o gpurun_out/prof_memops python tools/ops_once.py > gpurun_out/ncu_memops.log 2>&1; echo "ncu memops rc=$?"; tail -2 gpurun_out/ncu_memops.log
