#!/bin/bash
# fused vs two-kernel MixedOp forward: step time at B=96 and per-kernel device times at B=96 / B=8192
mkdir -p gpurun_out
for f in auto 0; do
  BMNAS_FUSED_MIXED=$f timeout 400 python bench.py --steps 300 --warmup 10 --no-cpu --roofline-batch 8192 --profile-kernels > gpurun_out/perf_fused_$f.log 2>&1
  echo "== BMNAS_FUSED_MIXED=$f"
  grep "^{" gpurun_out/perf_fused_$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches/step',d['launches_per_step'])
for k,v in d['roofline_large_batch']['kernels'].items(): print('  B8192', k, v)"
  grep "graph replay" gpurun_out/perf_fused_$f.log
  grep "^fwd" gpurun_out/perf_fused_$f.log | head -8
done
