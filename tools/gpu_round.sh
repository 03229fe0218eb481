#!/bin/bash
# one GPU call: parity tests, smoke, bench (+ per-kernel device times), reference arm, other configs,
# ncu launch list, ncu --set full captures of the hot kernels (B=96 through bench, B=8192 through the plan driver)
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 --profile-kernels > gpurun_out/bench_full.log 2>&1
grep "^{" gpurun_out/bench_full.log > gpurun_out/bench.log
cut -c1-300 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.log 2>&1
cut -c1-200 gpurun_out/bench_reference.log | tail -1
for cfg in mmimdb ego ego_large; do
  timeout 300 python bench.py --config $cfg --steps 50 --warmup 5 --no-cpu --roofline-batch 0 2>&1 | grep "^{" > gpurun_out/bench_$cfg.log
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$cfg.log').read());print('$cfg',d['value'],'samples/s',d['ms_per_step'],'ms/step e2e',d['e2e']['value'])" 2>&1 | tail -1
done
timeout 600 bash tools/ncu_launch_list.sh > gpurun_out/launch_list.txt 2>&1
head -16 gpurun_out/launch_list.txt
cap() {  # name, mangled-name regex, extra bench args
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 12 -c 1 \
    -o gpurun_out/prof_$1 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu --roofline-batch 0 $3 > gpurun_out/ncu_$1.log 2>&1
}
capL() {  # name, mangled-name regex, plan-driver filter: the B=8192 plan, every call launched stand-alone
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 2 -c 1 \
    -o gpurun_out/prof_$1 -f python tools/plan_kernels.py 8192 $3 eager > gpurun_out/ncu_$1.log 2>&1
}
cap node_fwd_B96 "k_node_fwdILi4" ""
cap node_bwd_B96 "k_node_bwdILi4" ""
cap sg_fwd_B96 k_sgILi0E ""
cap sg_dgrad_B96 k_sgILi1E ""
cap wgrad_B96 k_sgw ""
cap mix_bwd_B96 k_mix_bwd ""
cap ln_bwd_B96 k_ln_bwd ""
capL node_fwd_warp_B8192 k_node_fwd_warp node_fwd
capL node_bwd_warp_B8192 k_node_bwd_warp node_bwd
capL panel_fwd_B8192 k_gemm_panelILi0E conv_fwd
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
