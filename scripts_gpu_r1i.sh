#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
for B in 256 512 1024 8192; do for v in 1 2; do
  echo "B=$B variant=$v: $(BMNAS_NODE_VARIANT=$v timeout 300 python scripts_dbg_large.py $B node_ graph 2>&1 | grep '^fwd\|^bwd' | sed -n '1p;5p' | awk '{print $3, $(NF-1)}' | tr '\n' ' ')"
done; done
