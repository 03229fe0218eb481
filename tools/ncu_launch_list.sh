#!/bin/bash
# launch list (per-kernel device time) of one eager search step, then a full capture of the node kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 260 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 6 --warmup 3 --no-graphs --no-cpu > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches.csv', errors='ignore')))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = d['Kernel Name'][:70]
    v = float(d['Metric Value'].replace(',', ''))
    unit = d['Metric Unit']
    if unit == 'ns': v /= 1000.0
    elif unit == 'ms': v *= 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('total us', round(tot, 1))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{a[1]:9.1f} us  {100*a[1]/tot:5.1f}%  n={a[0]:4d}  avg={a[1]/a[0]:7.2f}  {k}')
PY
