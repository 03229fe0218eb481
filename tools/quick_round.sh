#!/bin/bash
# one short GPU call: all -m gpu tests (no -x: see every failure), then a quick bench with per-kernel times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/tests.log
tail -25 gpurun_out/tests.log
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu --roofline-batch 0 --profile-kernels 2>&1 > gpurun_out/bench_q.log
grep "^{" gpurun_out/bench_q.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches/step',d['launches_per_step'])"
grep "graph replay" gpurun_out/bench_q.log
