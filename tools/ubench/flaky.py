"""determinism stress of the tiny FusionNode used by test_gradcheck_dropout_mask_reuse"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200')); sys.path.insert(0, ROOT)
import torch
from models.search.darts.model_search import FusionNetwork  # noqa (import order)
from models.search.darts.node_search import FusionNode
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
args = types.SimpleNamespace(C=8, L=4, drpt=0.0, num_input_nodes=2, node_steps=2, node_multiplier=2)
node = FusionNode(2, 2, args).to(dev).train()
with torch.no_grad():
    node.gammas.copy_(torch.randn(2, 4, generator=g)); node.betas.copy_(torch.randn(5, 2, generator=g))
for m in node.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
x = torch.randn(6, 8, 4, generator=g).to(dev); w = torch.randn(6, 8, 4, generator=g).to(dev)
outs, gg, gb, gw = [], [], [], []
for it in range(300):
    with torch.no_grad():
        o = node(x, x).clone()
    outs.append(o)
ref = outs[0]
bad = [i for i, o in enumerate(outs) if not torch.equal(o, ref)]
print('forward: %d of %d runs differ from run 0; max abs diff %.3e' % (len(bad), len(outs), max([(o - ref).abs().max().item() for o in outs])), bad[:10])
prog = [r for r in node._bm_cache.values()][0].prog
names = [c.name for c in prog.fwd]
print(names)
for it in range(100):
    for p_ in list(node.parameters()) + [node.gammas, node.betas]:
        p_.grad = None
    out = node(x, x); (out * w).sum().backward()
    gg.append(node.gammas.grad.clone()); gb.append(node.betas.grad.clone())
    gw.append(torch.cat([p_.grad.flatten() for p_ in node.parameters() if p_.grad is not None]).clone())
for name, lst in (('gammas.grad', gg), ('betas.grad', gb), ('weights.grad', gw)):
    r = lst[0]
    bad = [i for i, o in enumerate(lst) if not torch.equal(o, r)]
    print('%s: %d of %d differ; max abs diff %.3e (|ref| max %.3e)' % (name, len(bad), len(lst), max([(o - r).abs().max().item() for o in lst]), r.abs().max().item()), bad[:10])
