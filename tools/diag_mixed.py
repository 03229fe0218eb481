"""diagnostic: NodeMixedOp module fwd/bwd vs oracle for fused on/off, injected masks, several batch sizes"""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import test_gpu_mixed as T
import gpu_util as U
from bmnas import program, native as N
L = 8
for B in (2048, 2496, 2500, 640, 800):
    for fused in ('1', '0'):
        for variant in (0, 1):
            program.FUSED_MIXED = fused
            N.lib().bmnas_set_node_variant(variant)
            mod = T._mixed(L).to(U.DEV).train()
            sd0 = {k: v.clone() for k, v in mod.state_dict().items()}
            g = torch.Generator().manual_seed(1)
            x = torch.randn(B, T.C, L, generator=g).to(U.DEV).requires_grad_(True)
            w = torch.softmax(torch.randn(4, generator=g), -1).to(U.DEV).requires_grad_(True)
            go = torch.randn(B, T.C, L, generator=g).to(U.DEV)
            masks = T._masks(mod, B, L, 2)
            U.inject_masks(mod, masks)
            out = mod(x, x, w); out.backward(go); torch.cuda.synchronize()
            ref = T._oracle(T._restore(T._mixed(L), sd0), x, w, go, masks, True, L, 0.2)
            e = lambda a, b: ((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()
            names = [c.name for r in mod._bm_cache.values() for c in r.prog.fwd + r.prog.bwd]
            worst = max((e(p.grad, ref[3]['mix.' + k]), k) for k, p in mod.named_parameters() if not k.endswith('conv.bias'))
            print(f'B={B} fused={fused} variant={variant}: out {e(out, ref[0]):.1e} gx {e(x.grad, ref[1]):.1e} gw {e(w.grad, ref[2]):.1e} worst param {worst[0]:.1e} {worst[1]}  kernels {sorted(set(names))}')
