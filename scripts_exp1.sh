#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do for gm in 0 1 2; do
  echo "== PDL=$pdl GEMM_MODE=$gm"
  BMNAS_PDL=$pdl BMNAS_GEMM_MODE=$gm python bench.py --steps 200 --warmup 10 --no-cpu 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['launches_per_step'])"
done; done
for B in 1024 8192; do
  echo "== batch $B"
  python bench.py --steps 20 --warmup 3 --no-cpu --batch $B --profile-kernels 2>&1 | tee gpurun_out/bench_B$B.log | grep -v "^{" | head -60
  grep "^{" gpurun_out/bench_B$B.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline'])"
done
