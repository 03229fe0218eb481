#!/bin/bash
# compute-sanitizer passes over the parity tests that exercise the warp-per-sample node kernels and the pool kernels
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider \
  -k "node_warp and (primitives or mixed5 or reshape_layers_golden or search_fwd_bwd_golden)" > gpurun_out/memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/memcheck.log | tail -6
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider \
  -k "node_warp and (primitives or mixed5)" > gpurun_out/racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" gpurun_out/racecheck.log | tail -8
