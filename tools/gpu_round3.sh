#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/mem_probe.py 4 > gpurun_out/mem_probe.log 2>&1; tail -45 gpurun_out/mem_probe.log
