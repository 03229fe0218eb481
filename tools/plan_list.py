"""print the launch plans of a search half step (validate-only, no GPU): python tools/plan_list.py [config] [arch|weights]"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200')):
    sys.path.insert(0, p)
import torch
import bench
from bmnas import native as N, runtime as rt
from bmnas.nn import SearchHead, CrossEntropyLoss, BCEWithLogitsLoss
N.set_validate_only(True)
c = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'ntu'])
mode = sys.argv[2] if len(sys.argv) > 2 else 'arch'
a = types.SimpleNamespace(**{k: c[k] for k in ('C', 'L', 'num_input_nodes', 'steps', 'multiplier', 'node_steps', 'node_multiplier', 'drpt')},
                          weight_decay=c['weight_decay'])
crit = CrossEntropyLoss() if c['loss'] == 'ce' else BCEWithLogitsLoss()
head = SearchHead(a, c['classes'], criterion=crit)
head.train()
feats = [torch.randn(c['B'], c['C'], c['L']) for _ in range(c['num_input_nodes'])]
y = torch.randint(0, c['classes'], (c['B'],)) if c['loss'] == 'ce' else torch.rand(c['B'], c['classes'])
with rt.grad_mode(mode), rt.static_io():
    res = head.loss_fused(feats, y, crit)
    (res[0] if res else crit(head(feats), y)).backward()
r = [r for k, r in head.fusion_net._bm_cache.items() if mode in k][0]
for ph, calls in (('fwd', r.prog._prep_calls + r.prog.fwd), ('bwd', r.prog.bwd)):
    for i, cl in enumerate(calls):
        st = cl.st
        dims = {k: getattr(st, k) for k in ('n', 'K', 'M', 'Ctot', 'n_src', 'mode', 'n_ops', 'n_chain') if hasattr(st, k)}
        extra = ''
        if cl.name == 'bmnas_mix_bwd':
            extra = ' gx=%d gw=%s gout2=%s' % (sum(1 for j in range(st.n) if st.gx[j]), bool(st.gw), bool(st.gout2))
        print('%s %2d %-22s %s %s%s' % (ph, i, cl.name, 'SIDE' if cl.side else '    ', dims, extra))
