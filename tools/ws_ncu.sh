#!/bin/bash
# ncu --set full of the warp-specialised GEMM kernels at B=8192 (plan driver) + in-kernel timeline + probe
mkdir -p gpurun_out
cap() { timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 2 -c 1 -o gpurun_out/prof_$1 -f python tools/plan_kernels.py 8192 $3 eager > gpurun_out/ncu_$1.log 2>&1; }
cap ws_dgrad_B8192 k_gemm_wsILi1E conv_dgrad
cap ws_fwd_B8192 k_gemm_wsILi0E conv_fwd
timeout 200 python tools/ws_timeline.py 8192 2>&1 | grep "^==\|stages\|done\|finalize"
PROBE_MODES=1 timeout 300 python tools/gemm_probe.py 8192 1024 2>&1 | grep "^B="
