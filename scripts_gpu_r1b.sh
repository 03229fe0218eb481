#!/bin/bash
# GPU call 1 of this session: tests, smoke, safe bench, and a hunt for the large-batch standalone-launch fault
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 10 --profile-kernels --roofline-batch 0 > gpurun_out/bench_safe_full.log 2>&1
grep "^{" gpurun_out/bench_safe_full.log > gpurun_out/bench_safe.log
cut -c1-600 gpurun_out/bench_safe.log
for B in 1024 8192; do
  for mode in eager graph; do
    timeout 300 python scripts_dbg_large.py $B "" $mode > gpurun_out/dbg_${B}_${mode}.log 2>&1
    echo "== dbg B=$B $mode: $(grep -c ' ok$\| us$' gpurun_out/dbg_${B}_${mode}.log) calls fine; last lines:"
    grep -v Warning gpurun_out/dbg_${B}_${mode}.log | grep "fwd \|bwd \|all ok\|Error\|error" | tail -3 | cut -c1-300
  done
done
# sanitizer on the first failing call (if any)
for f in gpurun_out/dbg_8192_eager.log gpurun_out/dbg_8192_graph.log gpurun_out/dbg_1024_eager.log gpurun_out/dbg_1024_graph.log; do
  if ! grep -q "all ok" $f; then
    name=$(grep "^fwd \|^bwd " $f | tail -1 | awk '{print $3}')
    B=$(echo $f | sed 's/.*dbg_\([0-9]*\)_.*/\1/')
    echo "== sanitizer on $name at B=$B"
    timeout 900 compute-sanitizer --print-limit 6 python scripts_dbg_large.py $B $name eager > gpurun_out/sanitizer.log 2>&1
    grep -v "^$" gpurun_out/sanitizer.log | grep -A14 "Invalid\|=== ERROR\|ERROR SUMMARY" | head -60
    break
  fi
done
