#!/bin/bash
# one GPU round: parity tests, smoke, bench (+ warm per-kernel profile)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/tests.log
tail -30 gpurun_out/tests.log
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 200 --warmup 10 --profile-kernels > gpurun_out/bench_full.log 2>&1
grep "^{" gpurun_out/bench_full.log > gpurun_out/bench.log
grep -v "^{" gpurun_out/bench_full.log | tail -50
