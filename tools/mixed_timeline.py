"""timeline of one bmnas_mixed_fwd launch (CTA 0 %globaltimer stamps): python tools/mixed_timeline.py [B]"""
import sys, os, types, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import test_gpu_mixed as T
import gpu_util as U
from bmnas import native as N
if os.environ.get('MODE'):
    N.lib().bmnas_set_gemm_mode(int(os.environ['MODE']))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
L = 8
mod = T._mixed(L).to(U.DEV).train()
x = torch.randn(B, T.C, L, device=U.DEV)
w = torch.softmax(torch.randn(4), -1).to(U.DEV)
for grad in (False, True):
    with torch.set_grad_enabled(grad):
        xx = x.clone().requires_grad_(grad)
        for _ in range(3):
            out = mod(xx, xx, w)
    torch.cuda.synchronize()
    runner = [r for r in mod._bm_cache.values() if r.prog.want_backward == grad][0]
    call = [c for c in runner.prog.fwd if c.name == 'bmnas_mixed_fwd'][0]
    ws_ptr = call.args[2].value
    ws = [t for t in runner.prog._keep if torch.is_tensor(t) and t.data_ptr() == ws_ptr][0]
    wi = ws.view(torch.int32)
    wi[3] = 1
    torch.cuda.synchronize()
    with torch.set_grad_enabled(grad):
        out = mod(xx, xx, w)
    torch.cuda.synchronize()
    off = (16 + 2 * 384 * 2 * 8) // 4
    tl = ws.view(torch.int32)[off:off + 32].cpu().view(torch.int64).tolist()
    names = ['entry', 'prologue done', 'prod stage0', 'prod last', 'mma b_full0', 'mma tile0 issued', 'epi t_full0', 'pre grid barrier',
             'post grid barrier', 'attn P ready', 'phase B start', 'epi done', 'teardown']
    print(f'B={B} grad={grad}  (us since entry)')
    for i, n in enumerate(names):
        if tl[i]:
            print(f'  {n:20s} {(tl[i] - tl[0]) / 1e3:8.2f}')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.set_grad_enabled(grad):
        e0.record()
        for _ in range(20):
            out = mod(xx, xx, w)
        e1.record(); torch.cuda.synchronize()
    print('  stream-launched avg per forward (incl. host):', e0.elapsed_time(e1) / 20 * 1e3, 'us')
