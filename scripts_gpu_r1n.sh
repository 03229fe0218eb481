#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -30 > gpurun_out/tests.log
tail -8 gpurun_out/tests.log
timeout 300 python bench_found.py --batches 1024,8192 --steps 20 2>&1 | grep "^{" | cut -c80-330
