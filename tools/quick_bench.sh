#!/bin/bash
# short bench run: value / e2e / roofline of the NTU search step
timeout 150 python bench.py --steps 50 --warmup 5 --no-cpu --roofline-batch 0 2>&1 | grep "^{" | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'], d['launches_per_step'])"
