#!/bin/bash
# round-2 compute-sanitizer passes: memcheck over the warp-specialised GEMMs (gemm_ws.cu, wgrad_ws.cu), the fused MixedOp
# kernels (mixed_small.cu, mixed_tc.cu), the fused head and the wide-CTA node / LayerNorm kernels; racecheck over the
# small fused kernels (shared-memory hazards).  tcgen05 / TMA kernels: racecheck does not model the async proxy, memcheck does
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mixed.py tests/test_gpu_head.py -q -x -p no:cacheprovider \
  -k "ws_engine or small or fused_mixed_vs_oracle or head" > gpurun_out/memcheck2.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/memcheck2.log | tail -6
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider \
  -k "shape_sweep and (C256-L16 or C192-L16-B7 or C64-L4)" > gpurun_out/memcheck3.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/memcheck3.log | tail -6
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_mixed.py tests/test_gpu_head.py -q -x -p no:cacheprovider \
  -k "small or head" > gpurun_out/racecheck2.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" gpurun_out/racecheck2.log | tail -8
