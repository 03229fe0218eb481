#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/tests.log
tail -12 gpurun_out/tests.log
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu --roofline-batch 0 --profile-kernels 2>&1 > gpurun_out/bench_q.log
grep "^{" gpurun_out/bench_q.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches/step',d['launches_per_step'])"
grep "conv_fwd\|conv_dgrad" gpurun_out/bench_q.log | head -6
