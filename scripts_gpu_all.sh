#!/bin/bash
# one GPU call: parity tests, smoke, bench (+ per-kernel device times), ncu launch list, ncu --set full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 200 --warmup 10 --profile-kernels > gpurun_out/bench_full.log 2>&1
grep "^{" gpurun_out/bench_full.log > gpurun_out/bench.log
cut -c1-400 gpurun_out/bench.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.log 2>&1
cut -c1-300 gpurun_out/bench_reference.log | tail -1
bash scripts_ncu_list.sh > gpurun_out/launch_list.txt 2>&1
head -24 gpurun_out/launch_list.txt
cap() {  # name, mangled-name regex, extra bench args
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 12 -c 1 \
    -o gpurun_out/prof_$1 -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu --roofline-batch 0 $3 > gpurun_out/ncu_$1.log 2>&1
}
cap node_fwd_B96 k_node_fwd ""
cap node_bwd_B96 k_node_bwd ""
cap sg_fwd_B96 k_sgILi0E ""
cap sg_dgrad_B96 k_sgILi1E ""
cap wgrad_B96 k_gemm_tcILi2E ""
cap mix_bwd_B96 k_mix_bwd ""
cap ln_bwd_B96 k_ln_bwd ""
cap node_fwd_B8192 k_node_fwd "--batch 8192"
cap node_bwd_B8192 k_node_bwd "--batch 8192"
cap panel_fwd_B8192 k_gemm_panelILi0E "--batch 8192"
ls -la gpurun_out/*.ncu-rep
