#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_default.log 2>&1
echo "rc=$?"; grep -v Warning gpurun_out/bench_default.log | tail -5 | cut -c1-1500
if ! grep -q '^{' gpurun_out/bench_default.log; then
  timeout 900 compute-sanitizer --print-limit 4 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/sanitizer_bench.log 2>&1
  grep -v "^$" gpurun_out/sanitizer_bench.log | grep -B2 -A16 "Invalid\|ERROR SUMMARY" | head -80
fi
