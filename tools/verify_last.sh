#!/bin/bash
timeout 100 python -m pytest tests/test_gpu_parity.py -q --tb=line -p no:cacheprovider -k "full_size_vs_oracle and (inner_only or deep_node)" 2>&1 | tail -4
bash tools/quick_bench.sh
