#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/tests.log
tail -4 gpurun_out/tests.log
BMNAS_NODE_VARIANT=2 timeout 300 python scripts_dbg_large.py 8192 node_ graph 2>&1 | grep "^fwd\|^bwd" | sed -n '1p;5p'
BMNAS_NODE_VARIANT=2 timeout 300 python scripts_dbg_large.py 2048 node_ graph 2>&1 | grep "^fwd\|^bwd" | sed -n '1p;5p'
BMNAS_NODE_VARIANT=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_node_bwd_warp -s 2 -c 1 \
  -o gpurun_out/prof_node_bwd_warp_B8192 -f python scripts_dbg_large.py 8192 node_bwd eager > gpurun_out/ncu_warp.log 2>&1
tail -1 gpurun_out/ncu_warp.log
