#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 10 --profile-kernels > gpurun_out/bench_full.log 2>&1
grep "^{" gpurun_out/bench_full.log > gpurun_out/bench.log
cut -c1-260 gpurun_out/bench.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | grep "^{" > gpurun_out/bench_reference.log
cut -c1-200 gpurun_out/bench_reference.log
timeout 300 bash tools/ncu_launch_list.sh > gpurun_out/launch_list.txt 2>&1
head -12 gpurun_out/launch_list.txt
