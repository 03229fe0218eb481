#!/bin/bash
# short validation call: full GPU suite, large-batch roofline bench, cold-L2 ncu captures of the B=8192 kernels
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 1200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu --no-configs > gpurun_out/bench_d.log 2>&1
grep "^{" gpurun_out/bench_d.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
b=d['roofline_large_batch']
for k,v in b['kernels'].items(): print('  B8192', k, v)
print('  mixedop', {k:(v['us'],v['hbm_frac']) for k,v in b['mixedop'].items()})"
capL() { timeout 300 env $4 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" -s 2 -c 1 -o gpurun_out/prof_$1 -f python tools/plan_kernels.py 8192 $3 eager > gpurun_out/ncu_$1.log 2>&1; }
capL mixed_fwd_tf32_B8192 k_mixed_fwd mixed_fwd ""
capL mixed_fwd_bf16_B8192 k_mixed_fwd mixed_fwd BMNAS_GEMM_MODE=3
capL node_bwd_warp_B8192 k_node_bwd_warp node_bwd ""
capL ws_dgrad_B8192 k_gemm_wsILi1E conv_dgrad ""
capL ws_fwd_B8192 k_gemm_wsILi0E conv_fwd ""
capL ws_wgrad_B8192 k_wgrad_ws conv_wgrad ""
capL ln_bwd_B8192 k_ln_bwd ln_bwd ""
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
