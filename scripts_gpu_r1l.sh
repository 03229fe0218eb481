#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench_found.py --batches 96,1024,8192,32768 --steps 20 > gpurun_out/found_sweep.log 2>&1
echo "rc=$?"; grep -v Warning gpurun_out/found_sweep.log | tail -8 | cut -c1-400
