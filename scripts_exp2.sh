#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:k_gemm_tcILi${m}E" -s 12 -c 1 -o gpurun_out/prof_kgemmtc${m}_B96 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu > gpurun_out/ncu_tc${m}.log 2>&1
  tail -2 gpurun_out/ncu_tc${m}.log
done
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:k_gemm_tcILi0E" -s 6 -c 1 -o gpurun_out/prof_kgemmtc0_B8192 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu --batch 8192 > gpurun_out/ncu_tc0_big.log 2>&1
ls -la gpurun_out/*.ncu-rep
