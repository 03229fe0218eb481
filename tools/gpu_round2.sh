#!/bin/bash
# round-2 GPU call: parity tests, smoke, bench (+ per-kernel device times), reference arm, ncu launch list,
# ncu --set full captures of the hot kernels (B=96 through bench, B=8192 through the plan driver)
#   tools/gpu_round2.sh [notests] [nocaps]
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
if [ "$1" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
  tail -3 gpurun_out/tests.log
else shift; fi
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 --profile-kernels > gpurun_out/bench_full.log 2>&1
grep "^{" gpurun_out/bench_full.log > gpurun_out/bench.log
cut -c1-300 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference_full.log 2>&1
grep "^{" gpurun_out/bench_reference_full.log > gpurun_out/bench_reference.log
cut -c1-200 gpurun_out/bench_reference.log | tail -1
timeout 600 bash tools/ncu_launch_list.sh > gpurun_out/launch_list.txt 2>&1
head -16 gpurun_out/launch_list.txt
[ "$1" == "nocaps" ] && exit 0
cap() {  # name, mangled-name regex, extra bench args
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 12 -c 1 \
    -o gpurun_out/prof_$1 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu --no-configs --roofline-batch 0 $3 > gpurun_out/ncu_$1.log 2>&1
}
capL() {  # name, mangled-name regex, plan-driver filter: the B=8192 plan, every call launched stand-alone; extra env
  timeout 300 env $4 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 2 -c 1 \
    -o gpurun_out/prof_$1 -f python tools/plan_kernels.py 8192 $3 eager > gpurun_out/ncu_$1.log 2>&1
}
cap mixed_small_B96 k_mixed_small ""
cap node_bwd_B96 "k_node_bwdILi4" ""
cap sg_dgrad_B96 k_sgILi1E ""
cap sg_fwd_B96 k_sgILi0E ""
cap wgrad_B96 k_sgw ""
cap head_B96 k_head ""
cap mix_bwd_B96 k_mix_bwd ""
cap ln_bwd_B96 k_ln_bwd ""
capL mixed_fwd_tf32_B8192 k_mixed_fwd mixed_fwd ""
capL mixed_fwd_bf16_B8192 k_mixed_fwd mixed_fwd BMNAS_GEMM_MODE=3
capL node_bwd_warp_B8192 k_node_bwd_warp node_bwd ""
capL ws_dgrad_B8192 k_gemm_wsILi1E conv_dgrad ""
capL ws_fwd_B8192 k_gemm_wsILi0E conv_fwd ""
capL ws_wgrad_B8192 k_wgrad_ws conv_wgrad ""
capL ln_bwd_B8192 k_ln_bwd ln_bwd ""
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
