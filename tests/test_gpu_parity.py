"""-m gpu parity tests: the CUDA path (through the C ABI) against
  (a) the golden fixtures computed by the real reference, and
  (b) the CPU oracle on seeded inputs at the BASELINE sizes.
fp32 tolerance (north_star): 1e-5 relative on outputs and gradients."""
import numpy as np
import pytest
import torch

from helpers import O, load, sub, cfg_of, arch_of, unpickle_genotype, geno_plain, assert_close
from helpers import close_vs_referee as _close_vs_referee
import gpu_util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[0, 2], ids=['node_auto', 'node_warp'])
def node_variant(request):
    """every parity test runs twice: with the default kernel choice (CTA per sample at these batch sizes) and with
    the warp-per-sample node kernels forced wherever the shape is eligible (bmnas_set_node_variant)"""
    from bmnas import native as N
    from bmnas import program
    lib = N.lib()
    lib.bmnas_set_node_variant(request.param)
    chain = program.CHAIN_NODE
    if request.param == 2:
        program.CHAIN_NODE = False      # chained inner mixes live in the CTA kernels: keep every op on the warp kernels here
    yield request.param
    program.CHAIN_NODE = chain
    lib.bmnas_set_node_variant(0)


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
    # node_steps=2 > node_multiplier=1: the first inner state feeds nothing but the next edge mix, so its only
    # upstream gradient arrives through the chained mix (gout NULL, gout2 set in bmnas_node_bwd)
    'inner_only': dict(cfg=O.Cfg(32, 8, 4, 2, 2, 2, 1, 0.2), B=12, classes=5, kind='ce'),
    'deep_node': dict(cfg=O.Cfg(32, 8, 4, 2, 2, 3, 2, 0.2), B=12, classes=5, kind='ce'),
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
    dbl = lambda t: t.double() if t.is_floating_point() else t
    P64 = {k: dbl(v) for k, v in P.items()}
    lv64, logits64, gw64, ga64, _ = _oracle_fb(cfg, P64, [a.double() for a in arch], [f.double() for f in feats],
                                               dbl(labels), masks, kind)
    out = head([f.to(U.DEV) for f in feats])
    loss = _loss_mod(kind)(out, labels.to(U.DEV))
    loss.backward()
    torch.cuda.synchronize()
    _close_vs_referee(out, logits, logits64, TOL, 'logits')
    _close_vs_referee(loss, lv, lv64, TOL, 'loss')
    for k, p in head.named_parameters():
        _close_vs_referee(p.grad, gw[k], gw64[k], GTOL, 'grad ' + k, atol=_bias_atol(k))
    for i, a in enumerate(head.arch_parameters()):
        _close_vs_referee(a.grad, ga[i], ga64[i], GTOL, f'garch{i}')
    sd = head.state_dict()
    for k in Pc:
        if 'running' in k or 'num_batches' in k:
            assert_close(sd[k], Pc[k], 1e-5, k)


SWEEP = [(C, L, B) for C in (64, 128, 192, 256) for L in (4, 8, 16) for B in (24,)] + [(256, 16, 96), (192, 16, 7)]


@pytest.mark.parametrize('C,L,B', SWEEP, ids=[f'C{c}-L{l}-B{b}' for c, l, b in SWEEP])
def test_shape_sweep_vs_oracle(C, L, B):
    """SURVEY 7.2b property sweep: the whole searchable head (forward, loss, every weight and architecture gradient,
    BatchNorm buffers) against the oracle over C x L, i.e. over every dispatch the shapes select -- fused small-batch
    MixedOp vs conv + node kernels (one co-resident wave or not), 256 / 512 / 1024-thread CTAs of the node and
    LayerNorm-tail kernels (C*L from 256 to 4096), FFMA vs tensor-core GEMM engines, ragged batch."""
    cfg = O.Cfg(C, L, 4, 2, 2, 2, 2, 0.1)
    ncls, kind = 7, 'ce'
    P = O.init_params(cfg, ncls, seed=11, prefix='cell')
    arch = O.init_arch(cfg, seed=11, scale=0.5)
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=12, loss=kind)
    head = U.build_head(cfg, ncls, P, arch)
    head.train()
    masks = U.random_masks(head, B, cfg.C, cfg.L, 13, cfg.drpt)
    U.inject_masks(head, masks)
    lv, logits, gw, ga, Pc = _oracle_fb(cfg, P, arch, feats, labels, masks, kind)
    dbl = lambda t: t.double() if t.is_floating_point() else t
    lv64, logits64, gw64, ga64, _ = _oracle_fb(cfg, {k: dbl(v) for k, v in P.items()}, [a.double() for a in arch],
                                               [f.double() for f in feats], dbl(labels), masks, kind)
    out = head([f.to(U.DEV) for f in feats])
    loss = _loss_mod(kind)(out, labels.to(U.DEV))
    loss.backward()
    torch.cuda.synchronize()
    _close_vs_referee(out, logits, logits64, TOL, 'logits')
    _close_vs_referee(loss, lv, lv64, TOL, 'loss')
    for k, p in head.named_parameters():
        _close_vs_referee(p.grad, gw[k], gw64[k], GTOL, 'grad ' + k, atol=_bias_atol(k))
    for i, a in enumerate(head.arch_parameters()):
        _close_vs_referee(a.grad, ga[i], ga64[i], GTOL, f'garch{i}')
    sd = head.state_dict()
    for k in Pc:
        if 'running' in k or 'num_batches' in k:
            assert_close(sd[k], Pc[k], 1e-5, k)


# ------------------------------------------------------------------ full search loop (Architect + FusedAdam + schedule)
def _static_masks(head, d, prefix):
    """injected masks as static device tensors (refilled per step, so the same pointers work under graphs)"""
    m0 = sub(d, prefix)
    U.inject_masks(head, m0)
    return {n: m.injected_mask for n, m in head.named_modules() if hasattr(m, 'injected_mask')}


@pytest.mark.parametrize('name', ['search_ntu_small', 'search_mmimdb_small', 'search_ego_small', 'search_deep_small'])
@pytest.mark.parametrize('graphs', [False, True])
def test_search_loop_golden(name, graphs):
    from bmnas.search import SearchStep
    d = load(name)
    cfg = cfg_of(d)
    kind = _kind(d)
    P = sub(d, 'sd0/')
    for k, v in sub(d, 'fb/sd/').items():     # the generator's single fwd/bwd advanced the BN buffers once
        P[k] = v.clone()
    head = U.build_head(cfg, int(d['num_classes']), P, arch_of(d, 'arch0/'))
    head.train()
    h = d['loop/hyper']
    B = int(d['B'])
    ss = SearchStep(head, _loss_mod(kind), B, int(d['num_classes']), loss_kind=kind, eta_max=h[0], eta_min=h[1],
                    Ti=h[2], Tm=h[3], nbpe=h[4], weight_decay=h[5], arch_lr=h[6], arch_wd=h[7], use_graphs=graphs)
    # one set of static mask tensors shared by the dev and train halves
    static = _static_masks(head, d, 'loop/0/mask_dev/')
    if graphs:
        ss.prepare(warmup=2, restore=True)
    for s in range(int(d['nsteps'])):
        dev_feats = torch.stack(arch_of(d, f'loop/{s}/dev_feat/'))
        trn_feats = torch.stack(arch_of(d, f'loop/{s}/train_feat/'))
        ss.load('dev', dev_feats, torch.from_numpy(d[f'loop/{s}/dev_labels']))
        ss.load('train', trn_feats, torch.from_numpy(d[f'loop/{s}/train_labels']))
        # arch half with the dev masks, weight half with the train masks
        for n, t in static.items():
            t.copy_(torch.from_numpy(d[f'loop/{s}/mask_dev/{n}']))
        ss._run_half('dev')
        ss.sched.step()
        ss.w_opt.set_lr(float(ss.sched.eta))
        for n, t in static.items():
            t.copy_(torch.from_numpy(d[f'loop/{s}/mask_train/{n}']))
        ss._run_half('train')
        torch.cuda.synchronize()
        assert_close(ss.loss['train'], d[f'loop/{s}/train_loss'], 1e-4, f'train loss step {s}')
        assert abs(ss.sched.eta - float(d[f'loop/{s}/lr'])) < 1e-12
        for i, a in enumerate(head.arch_parameters()):
            assert_close(a, d[f'loop/{s}/arch/{i}'], 1e-4, f'arch {i} step {s}')
        assert geno_plain(head.genotype()) == geno_plain(unpickle_genotype(d[f'loop/{s}/genotype'])), s
    sd = head.state_dict()
    for k, v in sub(d, 'loop/sd_final/').items():
        tol = 2e-2 if k.endswith('conv.bias') else 2e-4   # see test_oracle_golden.py on BN-fed conv biases
        assert_close(sd[k], v, tol, 'final ' + k)


# ------------------------------------------------------------------ in-kernel Philox dropout
def test_philox_dropout_statistics_and_consistency():
    cfg = O.Cfg(64, 8, 4, 2, 2, 2, 2, 0.2)
    B, ncls = 64, 10
    P = O.init_params(cfg, ncls, seed=1, prefix='cell')
    arch = O.init_arch(cfg, seed=1, scale=0.3)
    head = U.build_head(cfg, ncls, P, arch)
    head.train()
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=4)
    feats = [f.to(U.DEV) for f in feats]
    crit = _loss_mod('ce')
    outs = []
    for _ in range(3):
        outs.append(head(feats).detach().clone())
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2]), 'masks must change per forward'
    # forward/backward mask consistency: finite-difference check of dL/dgamma through one fixed rng step is
    # impossible from outside (the step advances per forward), so check the keep rate and determinism instead:
    # the stand-alone primitive with p=0.5 keeps ~half the elements and scales by 2
    import types
    from models.search.darts import node_operations as nops
    fc = nops.ConcatFC(32, types.SimpleNamespace(drpt=0.5)).to(U.DEV).train()
    x = torch.randn(256, 32, 8, device=U.DEV)
    y = torch.randn(256, 32, 8, device=U.DEV)
    o = fc(x, y)
    fc.eval()                         # eval uses running stats; compare the support only
    nz = (o != 0).float().mean().item()
    # ReLU zeroes ~half, dropout another half of the rest
    assert 0.2 < nz < 0.3, nz
    # gradient flows only through kept elements and with the same mask as the forward
    fc.train()
    x.requires_grad_(True)
    o = fc(x, y)
    go = torch.ones_like(o)
    o.backward(go)
    assert torch.isfinite(x.grad).all()
    assert x.grad.abs().sum() > 0


def test_gradcheck_dropout_mask_reuse():
    """d/dgamma by central differences with injected masks (fp32 forward differences on a tiny problem)"""
    import types
    from models.search.darts.node_search import FusionNode
    g = torch.Generator().manual_seed(0)
    args = types.SimpleNamespace(C=8, L=4, drpt=0.0, num_input_nodes=2, node_steps=2, node_multiplier=2)
    node = FusionNode(2, 2, args).to(U.DEV).train()
    with torch.no_grad():
        node.gammas.copy_(torch.randn(2, 4, generator=g))
        node.betas.copy_(torch.randn(5, 2, generator=g))
    for m in node.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                       # attention's hard-coded 0.1 off: deterministic function
    x = torch.randn(6, 8, 4, generator=g).to(U.DEV)
    w = torch.randn(6, 8, 4, generator=g).to(U.DEV)
    out = node(x, x)
    (out * w).sum().backward()
    ga = node.gammas.grad.clone()
    gb = node.betas.grad.clone()
    eps = 1e-2
    for t, gt_ in ((node.gammas, ga), (node.betas, gb)):
        num = torch.zeros_like(t)
        for idx in range(t.numel()):
            with torch.no_grad():
                t.view(-1)[idx] += eps
                lp = (node(x, x) * w).sum().item()
                t.view(-1)[idx] -= 2 * eps
                lm = (node(x, x) * w).sum().item()
                t.view(-1)[idx] += eps
            num.view(-1)[idx] = (lp - lm) / (2 * eps)
        assert_close(gt_, num, 5e-2, 'numeric grad', atol=2e-3)


# ------------------------------------------------------------------ input pipeline: prefetch() == load()
@pytest.mark.parametrize('graphs', [False, True])
def test_prefetch_pipeline_matches_serial_loads(graphs):
    """SearchStep.prefetch() (copy stream, batch i+1 in flight while step i computes) must feed every half step
    exactly the batch load() would have: the same loss trajectory and architecture (dropout off; the split
    reductions use float atomics, so the comparison is to 1e-5, far below the step-to-step differences)."""
    from bmnas.search import SearchStep
    from bmnas.nn import CrossEntropyLoss
    cfg = O.Cfg(32, 8, 4, 2, 2, 2, 2, 0.0)
    B, ncls, K = 16, 7, 6
    P = O.init_params(cfg, ncls, seed=11, prefix='cell')
    arch = O.init_arch(cfg, seed=11, scale=0.3)
    batches = []
    for i in range(2 * K):
        f, y = O.synthetic_batch(cfg, B, ncls, seed=100 + i)
        batches.append((torch.stack(f).pin_memory(), y.pin_memory()))

    def run(pipelined):
        head = U.build_head(cfg, ncls, P, arch)
        head.train()
        for m in head.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        ss = SearchStep(head, CrossEntropyLoss(), B, ncls, use_graphs=graphs)
        ss.load('dev', *batches[0]); ss.load('train', *batches[1])
        ss.prepare(warmup=2, restore=True)
        losses = []
        if pipelined:
            ss.prefetch('dev', *batches[0]); ss.prefetch('train', *batches[1])
        for i in range(K):
            if not pipelined:
                ss.load('dev', *batches[2 * i]); ss.load('train', *batches[2 * i + 1])
            la, lw = ss.step()
            if pipelined and i + 1 < K:
                ss.prefetch('dev', *batches[2 * i + 2]); ss.prefetch('train', *batches[2 * i + 3])
            losses.append((la.item(), lw.item()))
        torch.cuda.synchronize()
        return losses, [a.detach().cpu().clone() for a in head.arch_parameters()]

    l0, a0 = run(False)
    l1, a1 = run(True)
    for (xa, xw), (ya, yw) in zip(l0, l1):
        assert abs(xa - ya) <= 1e-5 * max(1.0, abs(xa)) and abs(xw - yw) <= 1e-5 * max(1.0, abs(xw)), (l0, l1)
    for x, y in zip(a0, a1):
        assert_close(y, x, 1e-4, 'arch after pipelined loop')
    assert len({round(v[1], 6) for v in l0}) > 1      # the batches really differ step to step


# ------------------------------------------------------------------ reshape layers upstream of the cells
@pytest.mark.parametrize('tag', ['ntu_5d', 'ntu_skel', 'ntu_vec', 'ragged_down', 'ragged_up', 'ego_5d', 'mm_map', 'mm_vec',
                                 'mm_small'])
def test_reshape_layers_golden(tag):
    """ReshapeInputLayer / ReshapeInputLayer_MMIMDB on the CUDA path (bmnas_pool_* + conv + node kernels) against the
    reference's own modules: output, input gradient through the adaptive max pool, parameter gradients, BatchNorm
    buffers, eval-mode output"""
    import types
    from models.auxiliary.aux_models import ReshapeInputLayer, ReshapeInputLayer_MMIMDB
    d = load('reshape')
    C, L, mm = (int(v) for v in d[f'{tag}/meta'])
    x = torch.from_numpy(d[f'{tag}/x']).to(U.DEV).requires_grad_(True)
    cls = ReshapeInputLayer_MMIMDB if mm else ReshapeInputLayer
    mod = cls(x.shape[1], C, L, types.SimpleNamespace(drpt=0.2))
    mod.load_state_dict({k[3:]: v for k, v in sub(d, f'{tag}/sd0/').items()}, strict=True)
    mod.to(U.DEV).train()
    U.inject_masks(mod, {k[3:]: v for k, v in sub(d, f'{tag}/mask/').items()})
    out = mod(x)
    out.backward(torch.from_numpy(d[f'{tag}/go']).to(U.DEV))
    torch.cuda.synchronize()
    assert_close(out, d[f'{tag}/out'], TOL, 'out')
    assert_close(x.grad, d[f'{tag}/gx'], GTOL, 'gx')
    for k, p in mod.named_parameters():
        assert_close(p.grad, d[f'{tag}/g/op.{k}'], GTOL, k, atol=(1e-4 if k.endswith('conv.bias') else 1e-7))
    sd = mod.state_dict()
    for k, v in sub(d, f'{tag}/sd1/').items():
        assert_close(sd[k[3:]], v, 1e-5, k)
    mod.eval()
    with torch.no_grad():
        ev = mod(x.detach())
    assert_close(ev, d[f'{tag}/eval_out'], TOL, 'eval')


@pytest.mark.parametrize('shape,C_in', [((96, 512, 8, 8, 8), 512), ((96, 2048), 2048), ((32, 1024, 8, 4, 4), 1024)])
def test_reshape_layer_full_size_vs_oracle(shape, C_in):
    """NTU-sized reshape layers (C_in up to 2048: the one large-K GEMM next to the path) against the CPU oracle"""
    import types
    from models.auxiliary.aux_models import ReshapeInputLayer
    C, L = 128, 8
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g)
    mod = ReshapeInputLayer(C_in, C, L, types.SimpleNamespace(drpt=0.2))
    P = {'op.' + k: v.detach().clone() for k, v in mod.state_dict().items()}
    mask = (torch.rand(shape[0], C, L, generator=g) >= 0.2).to(torch.uint8)
    go = torch.randn(shape[0], C, L, generator=g)
    names = O.trainable_names(P)
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    Pl = dict(P); Pl.update(leaves)
    o = O.reshape_input(x, Pl, 'op', L, {'op.dropout': mask}, True, 0.2)
    o.backward(go)
    mod.to(U.DEV).train()
    U.inject_masks(mod, {'dropout': mask})
    out = mod(x.to(U.DEV))
    out.backward(go.to(U.DEV))
    torch.cuda.synchronize()
    assert_close(out, o, TOL, 'out')
    for k, p in mod.named_parameters():
        # conv.bias feeds a train-mode BatchNorm: its gradient is analytically ZERO, what both sides compute is the rounding
        # noise of sum_n (a GV + b Z + c) over up to 6144 O(1) terms (~N * 2^-24 * |term|), whose value depends on the
        # summation order of the engine (FFMA split-K, tcgen05 split-K, MKL): the bound is a noise bound
        assert_close(p.grad, leaves['op.' + k].grad, GTOL, k, atol=(5e-4 if k.endswith('conv.bias') else 1e-7))
