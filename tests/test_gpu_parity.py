"""-m gpu parity tests: the CUDA path (through the C ABI) against
  (a) the golden fixtures computed by the real reference, and
  (b) the CPU oracle on seeded inputs at the BASELINE sizes.
fp32 tolerance (north_star): 1e-5 relative on outputs and gradients."""
import numpy as np
import pytest
import torch

from helpers import O, load, sub, cfg_of, arch_of, unpickle_genotype, geno_plain, assert_close
import gpu_util as U

pytestmark = pytest.mark.gpu
TOL = 1e-5
GTOL = 3e-5      # gradients: max-abs error relative to the largest entry of the tensor


def _bias_atol(k):
    # conv biases that feed a train-mode BatchNorm have an analytically zero gradient (rounding noise only)
    return 2e-5 if k.endswith('conv.bias') else 1e-7


def _loss_mod(kind):
    from bmnas.nn import CrossEntropyLoss, BCEWithLogitsLoss
    return CrossEntropyLoss() if kind == 'ce' else BCEWithLogitsLoss()


def _kind(d):
    return 'ce' if d['fb/labels'].ndim == 1 else 'bce'


@pytest.mark.parametrize('name', ['search_ntu_small', 'search_mmimdb_small', 'search_ego_small', 'search_deep_small'])
def test_search_fwd_bwd_golden(name):
    d = load(name)
    cfg = cfg_of(d)
    head = U.build_head(cfg, int(d['num_classes']), sub(d, 'sd0/'), arch_of(d, 'arch0/'))
    head.train()
    assert U.inject_masks(head, sub(d, 'fb/mask/')) > 0
    feats = [t.to(U.DEV).requires_grad_(True) for t in arch_of(d, 'fb/feat/')]
    labels = torch.from_numpy(d['fb/labels']).to(U.DEV)
    logits = head(feats)
    loss = _loss_mod(_kind(d))(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    assert_close(logits, d['fb/logits'], TOL, 'logits')
    assert_close(loss, d['fb/loss'], TOL, 'loss')
    for i, f in enumerate(feats):
        assert_close(f.grad, d[f'fb/gfeat/{i}'], GTOL, f'gfeat{i}')
    for k, p in head.named_parameters():
        assert p.grad is not None, k
        assert_close(p.grad, d['fb/g/' + k], GTOL, 'grad ' + k, atol=_bias_atol(k))
    for i, a in enumerate(head.arch_parameters()):
        assert_close(a.grad, d[f'fb/ga/{i}'], GTOL, f'garch{i}')
    sd = head.state_dict()
    for k, v in sub(d, 'fb/sd/').items():
        assert_close(sd[k], v, 1e-5, 'buffer ' + k)
    head.eval()
    with torch.no_grad():
        ev = head([f.detach() for f in feats])
    assert_close(ev, d['eval/logits'], TOL, 'eval logits')


@pytest.mark.parametrize('name', ['found_ntu_golden', 'found_mixed', 'found_nm1'])
def test_found_golden(name):
    d = load(name)
    cfg = cfg_of(d)
    gt = unpickle_genotype(d['genotype'])
    head = U.build_head(cfg, int(d['num_classes']), sub(d, 'sd0/'), genotype=gt)
    head.train()
    U.inject_masks(head, sub(d, 'fb/mask/'))
    feats = [t.to(U.DEV).requires_grad_(True) for t in arch_of(d, 'fb/feat/')]
    labels = torch.from_numpy(d['fb/labels']).to(U.DEV)
    logits = head(feats)
    loss = _loss_mod('ce')(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    assert_close(logits, d['fb/logits'], TOL, 'logits')
    for i, f in enumerate(feats):
        ref = d[f'fb/gfeat/{i}']
        if np.abs(ref).max() == 0:
            assert f.grad is None or float(f.grad.abs().max()) == 0.0
        else:
            assert_close(f.grad, ref, GTOL, f'gfeat{i}')
    for k, p in head.named_parameters():
        assert_close(p.grad, d['fb/g/' + k], GTOL, 'grad ' + k, atol=_bias_atol(k))
    sd = head.state_dict()
    for k, v in sub(d, 'fb/sd/').items():
        assert_close(sd[k], v, 1e-5, 'buffer ' + k)
    head.eval()
    with torch.no_grad():
        ev = head([f.detach() for f in feats])
    assert_close(ev, d['eval/logits'], TOL, 'eval logits')


@pytest.mark.parametrize('op', ['Sum', 'ScaleDotAttn', 'LinearGLU', 'ConcatFC', 'CatConvMish'])
@pytest.mark.parametrize('alias', [False, True])
def test_primitives(op, alias):
    import types
    from models.search.darts import node_operations as nops
    d = load('primitives')
    args = types.SimpleNamespace(C=16, L=8, drpt=0.2)
    mod = {'Sum': lambda: nops.Sum(), 'ScaleDotAttn': lambda: nops.ScaledDotAttn(16, 8),
           'LinearGLU': lambda: nops.LinearGLU(16, args), 'ConcatFC': lambda: nops.ConcatFC(16, args),
           'CatConvMish': lambda: nops.CatConvMish(16, args)}[op]()
    P0 = {k[3:]: v for k, v in sub(d, f'{op}/sd0/').items()}
    mod.load_state_dict(P0)
    mod.to(U.DEV).train()
    masks = {k[3:]: v for k, v in sub(d, f'{op}/mask/').items()}
    U.inject_masks(mod, masks)
    x = torch.from_numpy(d[f'{op}/x']).to(U.DEV).requires_grad_(True)
    y = x if alias else torch.from_numpy(d[f'{op}/y']).to(U.DEV).requires_grad_(True)
    go = torch.from_numpy(d[f'{op}/go']).to(U.DEV)
    out = mod(x, y)
    out.backward(go)
    torch.cuda.synchronize()
    if not alias:
        assert_close(out, d[f'{op}/out'], TOL, 'out')
        assert_close(x.grad, d[f'{op}/gx'], GTOL, 'gx')
        assert_close(y.grad, d[f'{op}/gy'], GTOL, 'gy')
        for k, p in mod.named_parameters():
            assert_close(p.grad, d[f'{op}/g/op.{k}'], GTOL, k, atol=_bias_atol(k))
        sd = mod.state_dict()
        for k, v in sub(d, f'{op}/sd1/').items():
            assert_close(sd[k[3:]], v, 1e-5, k)
        mod.eval()
        with torch.no_grad():
            ev = mod(x.detach(), y.detach())
        assert_close(ev, d[f'{op}/eval_out'], TOL, 'eval')
    else:   # x is y: check against the oracle (folded-weight path of the conv kernels)
        Pc = {k: v.clone() for k, v in sub(d, f'{op}/sd0/').items()}
        xc = torch.from_numpy(d[f'{op}/x']).requires_grad_(True)
        names = O.trainable_names(Pc)
        leaves = {k: Pc[k].clone().requires_grad_(True) for k in names}
        Pl = dict(Pc); Pl.update(leaves)
        o = O.step_op(op, xc, xc, Pl, 'op', sub(d, f'{op}/mask/'), True, 0.2)
        o.backward(torch.from_numpy(d[f'{op}/go']))
        assert_close(out, o, TOL, 'out (aliased)')
        assert_close(x.grad, xc.grad, GTOL, 'gx (aliased)')
        for k, p in mod.named_parameters():
            assert_close(p.grad, leaves['op.' + k].grad, GTOL, k, atol=_bias_atol(k))


def test_mixed5_with_catconvmish():
    import types
    from models.search.darts import node_operations as nops
    d = load('primitives')
    args = types.SimpleNamespace(C=16, L=8, drpt=0.2)
    nops.STEP_STEP_OPS['CatConvMish'] = lambda C, L, a: nops.CatConvMish(C, a)
    nops.STEP_STEP_PRIMITIVES.append('CatConvMish')
    try:
        mod = nops.NodeMixedOp(16, 8, args)
    finally:
        nops.STEP_STEP_PRIMITIVES.pop()
        del nops.STEP_STEP_OPS['CatConvMish']
    mod.load_state_dict({k[4:]: v for k, v in sub(d, 'Mixed5/sd0/').items()})
    mod.to(U.DEV).train()
    U.inject_masks(mod, {k[4:]: v for k, v in sub(d, 'Mixed5/mask/').items()})
    x = torch.from_numpy(d['Mixed5/x']).to(U.DEV).requires_grad_(True)
    y = torch.from_numpy(d['Mixed5/y']).to(U.DEV).requires_grad_(True)
    w = torch.from_numpy(d['Mixed5/w']).to(U.DEV).requires_grad_(True)
    out = mod(x, y, w)
    out.backward(torch.from_numpy(d['Mixed5/go']).to(U.DEV))
    assert_close(out, d['Mixed5/out'], TOL, 'out')
    assert_close(w.grad, d['Mixed5/gw'], GTOL, 'gw')
    assert_close(x.grad, d['Mixed5/gx'], GTOL, 'gx')
    assert_close(y.grad, d['Mixed5/gy'], GTOL, 'gy')
    for k, p in mod.named_parameters():
        assert_close(p.grad, d['Mixed5/g/mix.' + k], GTOL, k, atol=_bias_atol(k))


def test_edge_mixed_op_standalone():
    import types
    from models.search.darts.operations import FusionMixedOp
    op = FusionMixedOp(16, 8, types.SimpleNamespace(drpt=0.1)).to(U.DEV)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 16, 8, generator=g)
    w = torch.softmax(torch.randn(2, generator=g), -1)
    go = torch.randn(5, 16, 8, generator=g)
    xc, wc = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    oc = O.edge_mix([xc], wc[None])
    oc.backward(go)
    xg, wg = x.to(U.DEV).requires_grad_(True), w.to(U.DEV).requires_grad_(True)
    og = op(xg, wg)
    og.backward(go.to(U.DEV))
    assert_close(og, oc, TOL, 'out')
    assert_close(xg.grad, xc.grad, GTOL, 'gx')
    assert_close(wg.grad, wc.grad, GTOL, 'gw')


def _oracle_fb(cfg, P, arch, feats, labels, masks, kind, genotype=None):
    Pc = {k: v.clone() for k, v in P.items()}
    return O.loss_and_grads(feats, labels, arch, Pc, masks, cfg, loss=kind, genotype=genotype) + (Pc,)


CONFIGS = {
    # BASELINE configs at their real sizes
    'ntu': dict(cfg=O.Cfg(128, 8, 8, 2, 2, 2, 2, 0.2), B=96, classes=60, kind='ce'),
    'mmimdb': dict(cfg=O.Cfg(192, 16, 6, 2, 2, 1, 1, 0.1), B=32, classes=23, kind='bce'),
    'ego': dict(cfg=O.Cfg(128, 8, 8, 2, 2, 3, 3, 0.05), B=96, classes=83, kind='ce'),
    'ragged': dict(cfg=O.Cfg(40, 8, 5, 2, 2, 2, 2, 0.2), B=37, classes=11, kind='ce'),
}


@pytest.mark.parametrize('name', list(CONFIGS))
def test_full_size_vs_oracle(name):
    c = CONFIGS[name]
    cfg, B, ncls, kind = c['cfg'], c['B'], c['classes'], c['kind']
    P = O.init_params(cfg, ncls, seed=3, prefix='cell')
    arch = O.init_arch(cfg, seed=3, scale=0.5)
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=2, loss=kind)
    head = U.build_head(cfg, ncls, P, arch)
    head.train()
    masks = U.random_masks(head, B, cfg.C, cfg.L, 5, cfg.drpt)
    U.inject_masks(head, masks)
    lv, logits, gw, ga, Pc = _oracle_fb(cfg, P, arch, feats, labels, masks, kind)
    out = head([f.to(U.DEV) for f in feats])
    loss = _loss_mod(kind)(out, labels.to(U.DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert_close(out, logits, TOL, 'logits')
    assert_close(loss, lv, TOL, 'loss')
    for k, p in head.named_parameters():
        assert_close(p.grad, gw[k], GTOL, 'grad ' + k, atol=_bias_atol(k))
    for i, a in enumerate(head.arch_parameters()):
        assert_close(a.grad, ga[i], GTOL, f'garch{i}')
    sd = head.state_dict()
    for k in Pc:
        if 'running' in k or 'num_batches' in k:
            assert_close(sd[k], Pc[k], 1e-5, k)
