#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "reshape" 2>&1 | tail -40 > gpurun_out/tests_reshape.log
tail -25 gpurun_out/tests_reshape.log
