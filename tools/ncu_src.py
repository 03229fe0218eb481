"""top stall lines of an .ncu-rep source page: python tools/ncu_src.py file.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == 'Address':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
tot = sum(int(d['# Samples'] or 0) for d in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(d[s] or 0) for d in data) for s in stalls}
print('stall mix:', ', '.join(f'{k[6:]}={100*v/max(tot,1):.0f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, d in enumerate(data):
    d['i'] = i
for d in sorted(data, key=lambda d: -int(d['# Samples'] or 0))[:top]:
    why = sorted(((int(d[s] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{d['i']:5d} {int(d['# Samples']):6d} {100*int(d['# Samples'])/max(tot,1):5.1f}%  {d['Source'].strip()[:70]:70s} {why}")
