#!/bin/bash
mkdir -p gpurun_out
for k in "k_gemm_tc<0" "k_gemm_tc<1" "k_gemm_tc<2"; do
  n=$(echo $k | tr -d '<_' )
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 12 -c 1 -o gpurun_out/prof_${n}_B96 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu > /dev/null 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_gemm_tc<0" -s 6 -c 1 -o gpurun_out/prof_kgemmtc0_B8192 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu --batch 8192 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
