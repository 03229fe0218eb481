#!/bin/bash
mkdir -p gpurun_out
BMNAS_NODE_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_node_fwd_warp -s 4 -c 1 \
  -o gpurun_out/prof_node_fwd_warp_B8192 -f python scripts_dbg_large.py 8192 node_fwd eager > gpurun_out/ncu_warp.log 2>&1
tail -3 gpurun_out/ncu_warp.log
ls -la gpurun_out/*.ncu-rep
