"""SASS evidence per kernel: mnemonic counts (tcgen05 = UTC*MMA / LDTM / UTCBAR, TMA bulk = UBLKCP, mbarrier = SYNCS,
FFMA, Philox IMAD.HI, ...) of every __global__ function in bm-nas_b200/build/*.o, plus a short excerpt around the first
tensor-core instruction of the tcgen05 kernels.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'FFMA', 'HMMA', 'LDGSTS', 'LDG', 'STG', 'RED', 'ATOM',
        'LDS', 'STS', 'SHFL', 'MUFU', 'BAR', 'UCGABAR', 'ELECT', 'IMAD.HI', 'ACQBULK', 'CCTL']
print('# cuobjdump -sass of bm-nas_b200/build/*.o (sm_100a); one row per kernel: instruction count and mnemonic counts')
print('# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS')
excerpts = []
for obj in sorted(glob.glob(os.path.join(ROOT, 'bm-nas_b200', 'build', '*.o'))):
    sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    fn, body = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            fn = m.group(1)
            body[fn] = []
        elif fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            body[fn].append(line)
    print(f'\n== {os.path.basename(obj)}')
    for fn, lines in body.items():
        name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name)[:80]
        cnt = collections.Counter()
        for l in lines:
            ins = re.sub(r'/\*.*?\*/', '', l).strip().rstrip(';')
            ins = re.sub(r'^@!?U?P\d+\s+', '', ins)
            op = ins.split()[0] if ins.split() else ''
            for k in KEYS:
                if op == k or op.startswith(k + '.') or (k == 'UTCMMA' and re.match(r'UTC[A-Z]*MMA', op)):
                    cnt[k] += 1
                    break
        shown = ' '.join(f'{k}={cnt[k]}' for k in KEYS if cnt[k])
        print(f'  {name:<80s} n={len(lines):5d}  {shown}')
        tc = [i for i, l in enumerate(lines) if re.search(r'UTC[A-Z]*MMA', l)]
        if tc:
            i0 = max(0, tc[0] - 4)
            excerpts.append((name, [re.sub(r'\s+/\* 0x[0-9a-f]+ \*/', '', l).rstrip() for l in lines[i0:tc[0] + 8]]))
print('\n# excerpts: first tensor-core instruction of each tcgen05 kernel (with the descriptor set-up before it)')
for name, ls in excerpts:
    print(f'\n-- {name}')
    for l in ls:
        print(l)
