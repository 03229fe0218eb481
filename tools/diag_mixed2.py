import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import test_gpu_mixed as T
import gpu_util as U
from bmnas import program, native as N
L = 8; B = 2500
Z = {}
for fused in ('1', '0'):
    program.FUSED_MIXED = fused
    mod = T._mixed(L).to(U.DEV).train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T.C, L, generator=g).to(U.DEV).requires_grad_(True)
    w = torch.softmax(torch.randn(4, generator=g), -1).to(U.DEV).requires_grad_(True)
    masks = T._masks(mod, B, L, 2)
    U.inject_masks(mod, masks)
    go = torch.randn(B, T.C, L, generator=g).to(U.DEV)
    out = mod(x, x, w); out.backward(go); torch.cuda.synchronize()
    prog = list(mod._bm_cache.values())[0].prog
    zs = [t for t in prog._keep if torch.is_tensor(t) and tuple(t.shape) == (B, 384, L)]
    v384 = [t.clone() for t in prog._keep if torch.is_tensor(t) and tuple(t.shape) == (384,)]
    Z[fused] = (zs[0].clone(), out.detach().clone(), v384, [z.clone() for z in zs], x.grad.clone())
d = (Z['1'][0] - Z['0'][0]).abs()
print('Z max diff', d.max().item(), 'out max diff', (Z['1'][1] - Z['0'][1]).abs().max().item())
bad = (d > 1e-4).nonzero()
print('bad count', bad.shape[0])
if bad.shape[0]:
    bs = bad[:, 0].unique(); rows = bad[:, 1].unique(); ls = bad[:, 2].unique()
    print('bad samples', bs[:20].tolist(), '... n', bs.numel(), 'rows n', rows.numel(), rows[:10].tolist(), 'l', ls.tolist())
    b0 = bs[0].item()
    print('sample', b0, 'tile', b0 * L // 64, 'fused Z', Z['1'][0][b0, rows[0], :].tolist(), 'ref Z', Z['0'][0][b0, rows[0], :].tolist())

for i, (a, b) in enumerate(zip(Z['1'][2], Z['0'][2])):
    print('vec384', i, 'max diff', (a - b).abs().max().item(), 'max', b.abs().max().item())
for i, (a, b) in enumerate(zip(Z['1'][3], Z['0'][3])):
    dd = (a - b).abs()
    print('BxMxL buf', i, 'max diff', dd.max().item(), 'max', b.abs().max().item(), 'bad', (dd > 1e-3).sum().item())
    if (dd > 1e-3).any():
        bad = (dd > 1e-3).nonzero(); print('  bad samples', bad[:, 0].unique()[:12].tolist(), 'n', bad[:, 0].unique().numel(), 'rows', bad[:, 1].unique()[:8].tolist(), bad[:,1].unique().numel())
dd = (Z['1'][4] - Z['0'][4]).abs(); bad = (dd > 1e-3).nonzero()
print('gx max diff', dd.max().item(), 'bad samples', bad[:, 0].unique()[:12].tolist(), 'n', bad[:, 0].unique().numel())
