#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
for sp in 1; do
BMNAS_SPLIT_MIX_BWD=$sp timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu --roofline-batch 0 2>&1 | grep "^{" | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('split=$sp value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches/step',d['launches_per_step'])"
done
