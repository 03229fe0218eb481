"""executed warp-instructions of an .ncu-rep kernel bucketed by CUDA source line: joins the SASS source page of
the report (per-instruction 'Instructions Executed', samples) with `nvdisasm --print-line-info` of the cubin by
instruction order.   python tools/ncu_lines.py report.ncu-rep build/obj.o mangled_kernel_substring [top]"""
import csv, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = None, []
for r in rows:
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for l in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
    if m:
        infn = kern in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.search(r'inlined at "[^"]+", line (\d+)', m.group(3))
        cur = (os.path.basename(m.group(1)), int(m.group(2)), int(inl.group(1)) if inl else None); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', l):
        lines.append(cur)
print('sass instructions: report', len(data), 'disasm', len(lines))
n = min(len(data), len(lines))
agg = {}
tot = 0
for d, ln in zip(data[:n], lines[:n]):
    e = int(d['Instructions Executed'] or 0); s = int(d['# Samples'] or 0)
    tot += e
    a = agg.setdefault(ln, [0, 0]); a[0] += e; a[1] += s
ts = sum(a[1] for a in agg.values())
print('total executed warp instructions', tot, 'samples', ts)
src = {}
for (f, ln, inl), (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get("BY_SAMPLES") else 0])[:top]:
    if f not in src:
        p = os.path.join('bm-nas_b200/csrc', f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ''
    print(f'{100*e/max(tot,1):5.1f}% inst {100*s/max(ts,1):5.1f}% smp  {f}:{ln}' + (f' <-{inl}' if inl else '') + f'  {text}')
