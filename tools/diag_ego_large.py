"""diagnostic: per-tensor gradient errors of the ego_large Philox plan for both node variants"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
from helpers import O
import gpu_util as U
from bmnas import native as N, program
from bmnas.nn import CrossEntropyLoss
cfg = O.Cfg(256, 16, 8, 4, 4, 3, 3, 0.05); B = 96; ncls = 83
for variant, chain in ((0, True), (0, False), (2, False)):
    N.lib().bmnas_set_node_variant(variant); program.CHAIN_NODE = chain
    P = O.init_params(cfg, ncls, seed=3, prefix='cell'); arch = O.init_arch(cfg, seed=3, scale=0.5)
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=2)
    head = U.build_head(cfg, ncls, P, arch); head.train()
    out = head([f.to(U.DEV) for f in feats]); loss = CrossEntropyLoss()(out, labels.to(U.DEV)); loss.backward()
    torch.cuda.synchronize()
    masks = U.philox_masks(head, B, U.training_program(head))
    Pc = {k: v.clone() for k, v in P.items()}
    lv, logits, gw, ga = O.loss_and_grads(feats, labels, arch, Pc, masks, cfg)
    dbl = lambda t: t.double() if t.is_floating_point() else t
    lv64, l64, gw64, ga64 = O.loss_and_grads([f.double() for f in feats], labels, [a.double() for a in arch], {k: dbl(v) for k, v in P.items()}, masks, cfg)
    rows = []
    for k, p in head.named_parameters():
        r64 = gw64[k]; e = (p.grad.double().cpu() - r64).abs().max().item(); e32 = (gw[k].double() - r64).abs().max().item()
        rows.append((e / max(r64.abs().max().item(), 1e-30), k, e, e32, r64.abs().max().item()))
    rows.sort(reverse=True)
    print('variant', variant, 'chain', chain, 'loss err', abs(loss.item() - lv64.item()))
    for r in rows[:6]:
        print('  rel %.2e  %s  err %.2e  cpu32 err %.2e  max %.2e' % r)
