"""timeline of one bmnas_mixed_small_fwd launch (CTA (0,0) %globaltimer stamps): python tools/small_timeline.py [B] [L]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import test_gpu_mixed as T
import gpu_util as U
from bmnas import program
program.FUSED_MIXED = '0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mod = T._mixed(L).to(U.DEV).train()
x = torch.randn(B, T.C, L, device=U.DEV)
w = torch.softmax(torch.randn(4), -1).to(U.DEV)
xx = x.clone().requires_grad_(True)
for _ in range(3):
    out = mod(xx, xx, w)
torch.cuda.synchronize()
runner = [r for r in mod._bm_cache.values() if r.prog.want_backward][0]
call = [c for c in runner.prog.fwd if c.name == 'bmnas_mixed_small_fwd'][0]
ws_ptr = call.args[2].value
ws = [t for t in runner.prog._keep if torch.is_tensor(t) and t.data_ptr() == ws_ptr][0]
ws.view(torch.int32)[2] = 1
torch.cuda.synchronize()
names = ['entry', 'pdl_wait passed', 'x tile landed', 'GEMM done', 'Z + BN partials out (G)', 'S + softmax done (A)', 'O + dropout done (A)',
         'LN partials out (A)', 'grid barrier passed', 'statistics merged', 'epilogue done']
for rep in range(3):
    out = mod(xx, xx, w)
    torch.cuda.synchronize()
    tl = ws.view(torch.int32)[4:4 + 32].cpu().view(torch.int64).tolist()
    print(f'B={B} L={L} rep {rep} (us since entry)')
    for i, n in enumerate(names):
        if tl[i]:
            print(f'  {n:28s} {(tl[i] - tl[0]) / 1e3:8.2f}')
