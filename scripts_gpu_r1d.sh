#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/tests.log
tail -15 gpurun_out/tests.log
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu > gpurun_out/bench_default.log 2>&1
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_default.log') if l.startswith('{')][0])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'serial',d['e2e'].get('serial_value'))
    for k,v in d['roofline_large_batch']['kernels'].items(): print(k,v)
except Exception as e:
    print('bench failed',e); print(open('gpurun_out/bench_default.log').read()[-3000:])
PY
for v in 1 2; do
  BMNAS_NODE_VARIANT=$v timeout 300 python scripts_dbg_large.py 8192 node_fwd graph 2>&1 | grep "^fwd" | head -2
  BMNAS_NODE_VARIANT=$v timeout 300 python scripts_dbg_large.py 1024 node_fwd graph 2>&1 | grep "^fwd" | head -1
done
