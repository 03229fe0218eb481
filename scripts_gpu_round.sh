#!/bin/bash
# one GPU round: parity tests, smoke, bench, ncu launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/tests.log
tail -40 gpurun_out/tests.log
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 200 --warmup 10 2>&1 | tail -5 | tee gpurun_out/bench.log
python bench.py --steps 100 --warmup 5 --no-graphs --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_nograph.log
