#!/bin/bash
# one GPU call: parity tests, smoke, bench (+ warm per-kernel profile), ncu launch list, ncu full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 200 --warmup 10 --profile-kernels > gpurun_out/bench_full.log 2>&1
grep "^{" gpurun_out/bench_full.log > gpurun_out/bench.log
cat gpurun_out/bench.log
bash scripts_ncu_list.sh > gpurun_out/launch_list.txt 2>&1
head -30 gpurun_out/launch_list.txt
for k in k_conv_fwd k_conv_dgrad k_conv_wgrad k_node_bwd k_node_fwd k_ln_bwd k_mix_bwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
