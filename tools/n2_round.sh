#!/bin/bash
# 2-GPU call: bench (peer-memory step and NCCL step) + rank-0 kernel timeline of the data-parallel step
mkdir -p gpurun_out
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --no-configs --roofline-batch 0 > gpurun_out/bench_n2.log 2>&1
grep "^{" gpurun_out/bench_n2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('peer', d['value'], d['ms_per_step'], d['e2e']['value'], d['dp_check'])" || tail -5 gpurun_out/bench_n2.log
run 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --no-configs --roofline-batch 0 --nccl > gpurun_out/bench_n2_nccl.log 2>&1
grep "^{" gpurun_out/bench_n2_nccl.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('nccl', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/bench_n2_nccl.log
run 29513 tools/timeline.py ntu 2>&1 | grep -v Warn > gpurun_out/timeline_n2.txt; grep -n "k_dp_adam\|one step\|k_adam" gpurun_out/timeline_n2.txt | head -8; tail -22 gpurun_out/timeline_n2.txt
run 29514 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu --no-configs --roofline-batch 0 --scaling strong 2>&1 | grep "^{" | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('strong', d['value'], d['ms_per_step'], d['e2e']['value'])"
