#!/bin/bash
# A/B of the warp-specialised GEMM engine: B=8192 per-kernel times and the Ego-large / NTU step, BMNAS_WS_GEMM=0 vs 1
mkdir -p gpurun_out
for w in 0 1; do
  echo "== BMNAS_WS_GEMM=$w"
  BMNAS_WS_GEMM=$w PROBE_MODES=1,2 timeout 300 python tools/gemm_probe.py 8192 1024 2>&1 | grep "^B="
  BMNAS_WS_GEMM=$w PROBE_MODES=1 PROBE_EGO=1 timeout 300 python tools/gemm_probe.py 96 2>&1 | grep "^B="
  BMNAS_WS_GEMM=$w timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu --no-configs --roofline-batch 8192 > gpurun_out/ws_ab_$w.log 2>&1
  grep "^{" gpurun_out/ws_ab_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ntu value',d['value'],'ms',d['ms_per_step'])
r=d['roofline_large_batch']
for k,v in r['kernels'].items(): print('  B8192', k, v)
print('  mixedop bwd', r['mixedop']['bwd'])" || tail -5 gpurun_out/ws_ab_$w.log
  BMNAS_WS_GEMM=$w timeout 300 python bench.py --config ego_large --steps 50 --warmup 5 --no-cpu --no-configs --roofline-batch 0 --profile-kernels > gpurun_out/ws_ego_$w.log 2>&1
  grep "^{" gpurun_out/ws_ego_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('ego_large value',d['value'],'ms',d['ms_per_step'])" || tail -5 gpurun_out/ws_ego_$w.log
  grep -E "^(fwd|bwd) " gpurun_out/ws_ego_$w.log | sort -k4 -n -r | head -12
done
