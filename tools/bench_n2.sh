#!/bin/bash
# N=2 data-parallel bench (NCCL all-reduce of the flat gradient arena inside the captured graphs)
mkdir -p gpurun_out
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.log 2>&1
echo "rc=$?"; grep "^{" gpurun_out/bench_n2.log | cut -c1-700; grep -v "^{" gpurun_out/bench_n2.log | grep -v Warning | tail -5
