#!/bin/bash
# engine crossover: small-N FFMA kernels (gemm_sg.cu) vs warp-specialised tcgen05 kernels (gemm_ws.cu / wgrad_ws.cu), NTU shapes
for B in 96 192 256 384 512 768; do
  echo "== B=$B  sg (fp32 FFMA)"
  PROBE_FMT=1 PROBE_MODES=1 BMNAS_TC_MIN_MACS_W=1000000000000 timeout 200 python tools/gemm_probe.py $B 2>&1 | grep "^B="
  echo "== B=$B  ws (tcgen05 3xTF32)"
  PROBE_FMT=0 PROBE_MODES=1 BMNAS_TC_MIN_MACS_W=0 timeout 200 python tools/gemm_probe.py $B 2>&1 | grep "^B="
done
