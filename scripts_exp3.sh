#!/bin/bash
for gm in 1 0; do
  echo "== GEMM_MODE=$gm"
  BMNAS_GEMM_MODE=$gm python bench.py --steps 100 --warmup 5 --no-cpu --profile-kernels 2>&1 | grep -v "^{" | tail -50
done
