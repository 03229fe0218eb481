#!/bin/bash
mkdir -p gpurun_out
cap() { timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 12 -c 1 -o gpurun_out/prof_$1 -f python bench.py --config ego_large --steps 3 --warmup 3 --no-graphs --no-cpu --no-configs --roofline-batch 0 > gpurun_out/ncu_$1.log 2>&1; }
cap node_fwd_egoL "k_node_fwdILi4" 
cap node_bwd_egoL "k_node_bwdILi4"
ls -la gpurun_out/prof_*egoL* | awk '{print $5, $9}'
