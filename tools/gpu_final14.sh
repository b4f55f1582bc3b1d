#!/bin/bash
# last sanity of the round: smoke() and a short default bench on the final tree
mkdir -p gpurun_out
timeout 100 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log | cut -c1-300
timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_final2.json | cut -c1-420
