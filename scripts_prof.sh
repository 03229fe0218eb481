#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 --no-cpu --profile-kernels > gpurun_out/kprof.log 2>&1
grep -v "^{" gpurun_out/kprof.log | tail -60
ncu --set full --clock-control none --import-source on -k regex:k_conv_dgrad -s 20 -c 2 -o gpurun_out/prof_dgrad -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_node_bwd -s 12 -c 2 -o gpurun_out/prof_nodebwd -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
