"""Tolerance diagnostics for tests/test_gpu_philox_large.py: run the named cases, record for every compared tensor our
distance and the CPU-fp32 reference's own distance to the fp64 referee (both relative to the tensor's max), print the
worst ones.  python tools/diag_tol.py [case:variant ...] [--reps N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import test_gpu_philox_large as T  # noqa: E402

REC = []


def rec(ours, ref32, ref64, tol, what, atol=0.0, knife=0, cpu_mult=3.0):
    ours = torch.as_tensor(ours).double().cpu()
    r32, r64 = ref32.double(), ref64.double()
    mx = r64.abs().max().item()
    d = (ours - r64).abs()
    cpu = (r32 - r64).abs().max().item()
    lim = max(tol * mx, cpu_mult * cpu) + atol
    nbad = 0
    if d.dim() >= 1 and d.shape[0] > 1:
        nbad = int((d.reshape(d.shape[0], -1).max(dim=1).values > lim).sum())
    REC.append((d.max().item() / max(lim, 1e-300), what, d.max().item(), lim, mx, cpu, nbad, knife, tuple(d.shape)))


T.close_vs_referee = rec
args = [a for a in sys.argv[1:] if ':' in a]
reps = int(sys.argv[sys.argv.index('--reps') + 1]) if '--reps' in sys.argv else 1
cases = args or ['ntu_B4096:0', 'mmimdb_B32:2', 'ego_large_B96:0', 'ego_large_B96:2']
for cs in cases:
    name, v = cs.split(':')
    for r in range(reps):
        REC.clear()
        try:
            T._set_variant(int(v))
            T._attempt(T.CASES[name], int(v), T.CASES[name].get('seed', 3) + 10 * r)
        except Exception as e:  # noqa: BLE001
            print('EXC', type(e).__name__, str(e)[:300])
        T._set_variant(0)
        REC.sort(key=lambda t: -t[0])
        print(f'== {name} variant {v} rep {r}: {len(REC)} tensors, {sum(1 for t in REC if t[0] > 1)} over the limit')
        for t in REC[:8]:
            print('   x%.2f %-75s err %.3e lim %.3e max %.3e cpu32 %.3e bad-slices %d (knife %d) %s' % t)
