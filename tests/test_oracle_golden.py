"""Pins oracle/bmnas_oracle.py against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import O, load, sub, cfg_of, arch_of, unpickle_genotype, geno_plain, assert_close

TOL = 2e-5   # fp32 vs fp32, different summation orders
SEARCH = ['search_ntu_small', 'search_mmimdb_small', 'search_ego_small', 'search_deep_small']
FOUND = ['found_ntu_golden', 'found_mixed', 'found_nm1']


def _loss_kind(d):
    return 'ce' if d['fb/labels'].ndim == 1 else 'bce'


@pytest.mark.parametrize('name', SEARCH)
def test_search_fwd_bwd(name):
    d = load(name)
    cfg = cfg_of(d)
    P = sub(d, 'sd0/')
    arch = arch_of(d, 'arch0/')
    feats = [t.requires_grad_(True) for t in arch_of(d, 'fb/feat/')]
    labels = torch.from_numpy(d['fb/labels'])
    masks = sub(d, 'fb/mask/')
    names = O.trainable_names(P)
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    Pl = dict(P); Pl.update(leaves)
    al = [a.clone().requires_grad_(True) for a in arch]
    logits = O.head_logits(feats, al, Pl, masks, True, cfg)
    kind = _loss_kind(d)
    lv = (torch.nn.functional.cross_entropy(logits, labels) if kind == 'ce'
          else torch.nn.functional.binary_cross_entropy_with_logits(logits, labels))
    lv.backward()
    assert_close(logits, d['fb/logits'], TOL, 'logits')
    assert_close(lv, d['fb/loss'], TOL, 'loss')
    for i, f in enumerate(feats):
        assert_close(f.grad, d[f'fb/gfeat/{i}'], 5e-5, f'gfeat{i}')
    for k in names:
        assert_close(leaves[k].grad, d['fb/g/' + k], 5e-5, 'grad ' + k, atol=(1e-5 if k.endswith('conv.bias') else 2e-7))
    for i, a in enumerate(al):
        assert_close(a.grad, d[f'fb/ga/{i}'], 5e-5, f'garch{i}')
    for k, v in sub(d, 'fb/sd/').items():
        assert_close(P[k], v, 1e-5, 'buffer ' + k)
    # eval forward with the updated running statistics
    with torch.no_grad():
        ev = O.head_logits([f.detach() for f in feats], arch, P, None, False, cfg)
    assert_close(ev, d['eval/logits'], TOL, 'eval logits')


@pytest.mark.parametrize('name', SEARCH)
def test_search_loop(name):
    d = load(name)
    cfg = cfg_of(d)
    P = sub(d, 'sd0/')
    arch = arch_of(d, 'arch0/')
    # the generator ran the single fwd/bwd first: BN buffers had advanced by one step
    for k, v in sub(d, 'fb/sd/').items():
        P[k] = v.clone()
    h = d['loop/hyper']
    st = O.SearchState(cfg, P, arch, eta_max=h[0], eta_min=h[1], Ti=h[2], Tm=h[3], nbpe=h[4],
                       weight_decay=h[5], arch_lr=h[6], arch_wd=h[7], loss=_loss_kind(d))
    for s in range(int(d['nsteps'])):
        dev = (arch_of(d, f'loop/{s}/dev_feat/'), torch.from_numpy(d[f'loop/{s}/dev_labels']))
        trn = (arch_of(d, f'loop/{s}/train_feat/'), torch.from_numpy(d[f'loop/{s}/train_labels']))
        la, lw = st.search_step(dev, trn, sub(d, f'loop/{s}/mask_dev/'), sub(d, f'loop/{s}/mask_train/'))
        assert_close(lw, d[f'loop/{s}/train_loss'], 1e-4, f'train loss step {s}')
        assert abs(st.sched.eta - float(d[f'loop/{s}/lr'])) < 1e-12
        for i, a in enumerate(st.arch):
            assert_close(a, d[f'loop/{s}/arch/{i}'], 1e-4, f'arch {i} step {s}')
        assert geno_plain(st.genotype()) == geno_plain(unpickle_genotype(d[f'loop/{s}/genotype']))
    for k, v in sub(d, 'loop/sd_final/').items():
        # conv biases feeding a train-mode BN have an analytically-zero gradient: Adam turns
        # rounding noise into +-lr steps there, so they are compared loosely.
        tol = 2e-2 if k.endswith('conv.bias') else 2e-4
        assert_close(P[k], v, tol, 'final ' + k)


@pytest.mark.parametrize('name', FOUND)
def test_found(name):
    d = load(name)
    cfg = cfg_of(d)
    gt = unpickle_genotype(d['genotype'])
    P = sub(d, 'sd0/')
    feats = [t.requires_grad_(True) for t in arch_of(d, 'fb/feat/')]
    labels = torch.from_numpy(d['fb/labels'])
    names = O.trainable_names(P)
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    Pl = dict(P); Pl.update(leaves)
    logits = O.head_logits(feats, None, Pl, sub(d, 'fb/mask/'), True, cfg, genotype=gt)
    lv = torch.nn.functional.cross_entropy(logits, labels)
    lv.backward()
    assert_close(logits, d['fb/logits'], TOL, 'logits')
    for i, f in enumerate(feats):
        g = f.grad if f.grad is not None else torch.zeros_like(f)
        assert_close(g, d[f'fb/gfeat/{i}'], 5e-5 if np.abs(d[f'fb/gfeat/{i}']).max() > 0 else 1.0, f'gfeat{i}')
    for k in names:
        assert_close(leaves[k].grad, d['fb/g/' + k], 5e-5, 'grad ' + k, atol=(1e-5 if k.endswith('conv.bias') else 2e-7))
    for k, v in sub(d, 'fb/sd/').items():
        assert_close(P[k], v, 1e-5, 'buffer ' + k)
    with torch.no_grad():
        ev = O.head_logits([f.detach() for f in feats], None, P, None, False, cfg, genotype=gt)
    assert_close(ev, d['eval/logits'], TOL, 'eval logits')
    # shapes/names of the found state_dict
    shp = O.param_shapes(cfg, int(d['num_classes']), genotype=gt)
    assert set(shp) == set(P)


def test_param_names_match_reference():
    for name in SEARCH:
        d = load(name)
        cfg = cfg_of(d)
        P = sub(d, 'sd0/')
        shp = O.param_shapes(cfg, int(d['num_classes']))
        assert list(shp) == list(P), name
        for k, (s, _) in shp.items():
            assert tuple(P[k].shape) == tuple(s), k
        assert [tuple(a.shape) for a in arch_of(d, 'arch0/')] == O.arch_shapes(cfg)


@pytest.mark.parametrize('op', ['Sum', 'ScaleDotAttn', 'LinearGLU', 'ConcatFC', 'CatConvMish'])
def test_primitives(op):
    d = load('primitives')
    P = sub(d, f'{op}/sd0/')
    x = torch.from_numpy(d[f'{op}/x']).requires_grad_(True)
    y = torch.from_numpy(d[f'{op}/y']).requires_grad_(True)
    names = O.trainable_names(P)
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    Pl = dict(P); Pl.update(leaves)
    o = O.step_op(op, x, y, Pl, 'op', sub(d, f'{op}/mask/'), True, 0.2)
    o.backward(torch.from_numpy(d[f'{op}/go']))
    assert_close(o, d[f'{op}/out'], TOL, 'out')
    assert_close(x.grad, d[f'{op}/gx'], 5e-5, 'gx')
    assert_close(y.grad, d[f'{op}/gy'], 5e-5, 'gy')
    for k in names:
        assert_close(leaves[k].grad, d[f'{op}/g/' + k], 5e-5, k, atol=(1e-5 if k.endswith('conv.bias') else 2e-7))
    for k, v in sub(d, f'{op}/sd1/').items():
        assert_close(P[k], v, 1e-5, k)
    with torch.no_grad():
        ev = O.step_op(op, x.detach(), y.detach(), P, 'op', None, False, 0.2)
    assert_close(ev, d[f'{op}/eval_out'], TOL, 'eval')


RESHAPE_TAGS = ['ntu_5d', 'ntu_skel', 'ntu_vec', 'ragged_down', 'ragged_up', 'ego_5d', 'mm_map', 'mm_vec', 'mm_small']


@pytest.mark.parametrize('tag', RESHAPE_TAGS)
def test_reshape_input_layer(tag):
    """oracle restatement of ReshapeInputLayer / ReshapeInputLayer_MMIMDB (aux_models.py:51-115) against the
    reference's own modules: output, input gradient (through the adaptive max pool), parameter gradients,
    BatchNorm buffers, eval-mode output"""
    d = load('reshape')
    C, L, mm = (int(v) for v in d[f'{tag}/meta'])
    P = sub(d, f'{tag}/sd0/')
    x = torch.from_numpy(d[f'{tag}/x']).requires_grad_(True)
    names = O.trainable_names(P)
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    Pl = dict(P); Pl.update(leaves)
    o = O.reshape_input(x, Pl, 'op', L, sub(d, f'{tag}/mask/'), True, 0.2, mmimdb=bool(mm))
    o.backward(torch.from_numpy(d[f'{tag}/go']))
    assert o.shape == (x.shape[0], C, L)
    assert_close(o, d[f'{tag}/out'], TOL, 'out')
    assert_close(x.grad, d[f'{tag}/gx'], 5e-5, 'gx')
    for k in names:
        # a conv bias feeding a train-mode BatchNorm has an analytically zero gradient: rounding noise on both sides
        assert_close(leaves[k].grad, d[f'{tag}/g/' + k], 5e-5, k, atol=(1e-4 if k.endswith('conv.bias') else 2e-7))
    for k, v in sub(d, f'{tag}/sd1/').items():
        assert_close(P[k], v, 1e-5, k)
    with torch.no_grad():
        ev = O.reshape_input(x.detach(), P, 'op', L, None, False, 0.2, mmimdb=bool(mm))
    assert_close(ev, d[f'{tag}/eval_out'], TOL, 'eval')


def test_mixed5_with_catconvmish():
    d = load('primitives')
    P = sub(d, 'Mixed5/sd0/')
    cfg = O.Cfg(16, 8, 2, 1, 1, 1, 1, 0.2, step_ops=O.STEP_STEP_PRIMITIVES + ['CatConvMish'])
    x = torch.from_numpy(d['Mixed5/x']).requires_grad_(True)
    y = torch.from_numpy(d['Mixed5/y']).requires_grad_(True)
    w = torch.from_numpy(d['Mixed5/w']).requires_grad_(True)
    o = O.node_mixed(x, y, w, P, 'mix', sub(d, 'Mixed5/mask/'), True, cfg)
    o.backward(torch.from_numpy(d['Mixed5/go']))
    assert_close(o, d['Mixed5/out'], TOL, 'out')
    assert_close(w.grad, d['Mixed5/gw'], 5e-5, 'gw')
    assert_close(x.grad, d['Mixed5/gx'], 5e-5, 'gx')


def test_genotypes():
    d = load('genotypes')
    for i in range(int(d['n'])):
        steps, mult, n_in, ns, nm = [int(v) for v in d[f'{i}/cfg']]
        cfg = O.Cfg(8, 4, n_in, steps, mult, ns, nm, 0.1)
        arch = arch_of(d, f'{i}/arch/')
        g = O.network_genotype(arch, cfg)
        ref = unpickle_genotype(d[f'{i}/pickle'])
        assert geno_plain(g) == geno_plain(ref), i
        assert str(g) == str(d[f'{i}/str']), i


def test_tie_break_vector():
    """all-equal alpha/beta/gamma (SURVEY 3.4 [probe])"""
    cfg = O.Cfg(8, 4, 8, 2, 2, 2, 2, 0.1)
    arch = [torch.zeros(s) for s in O.arch_shapes(cfg)]
    g = O.network_genotype(arch, cfg)
    assert g.edges == [('skip', 0), ('skip', 1), ('skip', 0), ('skip', 2)]
    assert all(s.inner_steps == ['Sum', 'Sum'] for s in g.steps)


def test_scheduler():
    d = load('scheduler')
    for i in range(int(d['n'])):
        h = d[f'{i}/hyper']
        sc = O.CosineRestartLR(h[0], h[1], h[2], h[3], h[4])
        lr = np.asarray([sc.step() for _ in range(400)])
        assert np.array_equal(lr, d[f'{i}/lr'])


@pytest.mark.parametrize('tag', ['weight', 'arch'])
def test_adam(tag):
    d = load('adam')
    lr, b1, b2, wd = [float(v) for v in d[f'{tag}/hyper']]
    ps = [torch.from_numpy(d[f'{tag}/p0/{j}'].copy()) for j in range(2)]
    st = {}
    for s in range(20):
        gs = [torch.from_numpy(d[f'{tag}/g/{s}/{j}']) for j in range(2)]
        O.adam_step(ps, gs, st, lr, (b1, b2), wd)
        for j in range(2):
            assert_close(ps[j], d[f'{tag}/p/{s}/{j}'], 1e-6, f'{tag} step {s} p{j}')
