"""CPU-only tests of the host side: the C-ABI binding, the drop-in surface (state_dict names, genotype
derivation, pickle format, LR schedule), launch-plan construction in validate-only mode, loud failures."""
import ctypes
import io
import os
import pickle
import re
import types

import numpy as np
import pytest
import torch

from helpers import O, ROOT, load, sub, cfg_of, arch_of, unpickle_genotype, geno_plain
import gpu_util as U

PKG = os.path.join(ROOT, 'bm-nas_b200')


@pytest.fixture
def validate_only():
    from bmnas import native as N
    N.set_validate_only(True)
    yield N
    N.set_validate_only(False)


def test_abi_loads_and_exports_every_declared_symbol():
    from bmnas import native as N
    lib = N.lib()
    hdr = open(os.path.join(ROOT, 'include', 'bmnas_b200.h')).read()
    declared = set(re.findall(r'\b(bmnas_\w+)\s*\(', hdr))
    assert len(declared) >= 20
    for fn in declared:
        assert hasattr(lib, fn), fn
    assert lib.bmnas_abi_version() == 1
    names = ['bmnas_mix_params', 'bmnas_conv_params', 'bmnas_node_params', 'bmnas_ln_params', 'bmnas_loss_params',
             'bmnas_adam_tensor', 'bmnas_adam_params']
    for i, n in enumerate(names):     # ctypes layout generated from the header == the compiler's layout
        assert ctypes.sizeof(N.STRUCTS[n]) == lib.bmnas_sizeof_params(i), n
    assert lib.bmnas_strerror(-1).decode().startswith('invalid')


def test_abi_rejects_bad_parameter_blocks(validate_only):
    N = validate_only
    lib = N.lib()
    st = N.bmnas_mix_params()
    assert lib.bmnas_mix_fwd(ctypes.byref(st), None) == N.BMNAS_EINVAL        # n = 0, null pointers
    cv = N.bmnas_conv_params()
    cv.B, cv.L, cv.K, cv.M, cv.n_src, cv.n_seg, cv.w_fold = 4, 8, 32, 16, 1, 1, 1
    cv.src_C[0], cv.seg_M[0] = 31, 16                                          # channels do not add up to K
    assert lib.bmnas_conv_fwd(ctypes.byref(cv), None) == N.BMNAS_EINVAL
    nd = N.bmnas_node_params()
    nd.B, nd.C, nd.L, nd.n_ops = 2, 8, 128, 1                                  # L > 64 unsupported
    assert lib.bmnas_node_fwd(ctypes.byref(nd), None) == N.BMNAS_EINVAL


def test_product_never_touches_the_oracle_or_reference():
    bad = []
    for dp, _, fs in os.walk(PKG):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r'import\s+oracle|from\s+oracle|bmnas_oracle|/root/reference', txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_no_cpu_fallback():
    from bmnas import native as N
    from bmnas.nn import SearchHead, CrossEntropyLoss
    cfg = O.Cfg(8, 4, 3, 2, 2, 1, 1, 0.1)
    head = SearchHead(U.args_of(cfg), 3)
    feats = [torch.randn(2, 8, 4) for _ in range(3)]
    with pytest.raises(N.NativeError):
        head(feats)
    with pytest.raises(N.NativeError):
        CrossEntropyLoss()(torch.randn(2, 3), torch.tensor([0, 1]))


def test_missing_library_fails_loudly(monkeypatch):
    from bmnas import native as N
    monkeypatch.setattr(N, '_lib', None)
    monkeypatch.setattr(N, 'LIB_PATH', '/nonexistent/libbmnas_b200.so')
    with pytest.raises(N.NativeError):
        N.lib()


@pytest.mark.parametrize('name', ['search_ntu_small', 'search_mmimdb_small', 'search_ego_small', 'search_deep_small'])
def test_state_dict_surface_matches_reference(name):
    d = load(name)
    cfg = cfg_of(d)
    from bmnas.nn import SearchHead
    head = SearchHead(U.args_of(cfg), int(d['num_classes']))
    ref = sub(d, 'sd0/')
    sd = head.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    head.load_state_dict(ref)                      # a reference-saved best_model.pt loads
    assert [tuple(a.shape) for a in head.arch_parameters()] == [tuple(a.shape) for a in arch_of(d, 'arch0/')]
    # architecture tensors are not Parameters / not in the state_dict, as in the reference
    assert not any('alpha' in k or 'beta' in k or 'gamma' in k for k in sd)


@pytest.mark.parametrize('name', ['found_ntu_golden', 'found_mixed', 'found_nm1'])
def test_found_state_dict_surface(name):
    d = load(name)
    cfg = cfg_of(d)
    from bmnas.nn import SearchHead
    gt = U.to_product_genotype(unpickle_genotype(d['genotype']))
    head = SearchHead(U.args_of(cfg), int(d['num_classes']), genotype=gt)
    ref = sub(d, 'sd0/')
    assert list(head.state_dict().keys()) == list(ref.keys())
    head.load_state_dict(ref)
    assert head.fusion_net.get_genotype() == gt


def test_genotype_derivation_matches_reference():
    from models.search.darts.model_search import FusionNetwork
    d = load('genotypes')
    for i in range(int(d['n'])):
        steps, mult, n_in, ns, nm = [int(v) for v in d[f'{i}/cfg']]
        args = types.SimpleNamespace(C=8, L=4, drpt=0.1, num_input_nodes=n_in, node_steps=ns, node_multiplier=nm)
        net = FusionNetwork(steps, mult, n_in, 2, args)
        with torch.no_grad():
            for a, b in zip(net.arch_parameters(), arch_of(d, f'{i}/arch/')):
                a.copy_(b)
        g = net.genotype()
        assert str(g) == str(d[f'{i}/str']), i
        # byte-identical pickle: same namedtuple names, field order and module path as the reference
        assert pickle.dumps(g) == np.asarray(d[f'{i}/pickle']).tobytes(), i
        assert pickle.loads(np.asarray(d[f'{i}/pickle']).tobytes()) == g


def test_tie_break_vector():
    from models.search.darts.model_search import FusionNetwork
    args = types.SimpleNamespace(C=8, L=4, drpt=0.1, num_input_nodes=8, node_steps=2, node_multiplier=2)
    net = FusionNetwork(2, 2, 8, 2, args)
    with torch.no_grad():
        for a in net.arch_parameters():
            a.zero_()
    g = net.genotype()
    assert g.edges == [('skip', 0), ('skip', 1), ('skip', 0), ('skip', 2)]
    assert all(s.inner_steps == ['Sum', 'Sum'] for s in g.steps)
    assert g.concat == [8, 9]


def test_legacy_genotype_upgrade():
    from models.search.darts.genotypes import Genotype, StepGenotype, upgrade_legacy
    old = Genotype(edges=[('skip', 2), ('skip', 4)], steps=[StepGenotype([('skip', 1), ('skip', 0)], ['cat_conv_relu'], [2])],
                   concat=[6])
    assert upgrade_legacy(old).steps[0].inner_steps == ['ConcatFC']


def test_scheduler_matches_reference():
    from models.auxiliary.scheduler import LRCosineAnnealingScheduler
    d = load('scheduler')
    for i in range(int(d['n'])):
        h = d[f'{i}/hyper']
        sc = LRCosineAnnealingScheduler(h[0], h[1], h[2], h[3], h[4])
        lr = np.asarray([sc.step() for _ in range(400)])
        assert np.array_equal(lr, d[f'{i}/lr'])

    class Opt:
        param_groups = [{'lr': 0.0}]
    sc.update_optimizer(Opt)
    assert Opt.param_groups[0]['lr'] == sc.eta


def test_arch_tensors_follow_module_to_device():
    from models.search.darts.model_search import FusionNetwork
    args = types.SimpleNamespace(C=8, L=4, drpt=0.1, num_input_nodes=3, node_steps=1, node_multiplier=1)
    net = FusionNetwork(2, 2, 3, 2, args)
    ids = [id(a) for a in net.arch_parameters()]
    opt = torch.optim.Adam(net.arch_parameters(), lr=3e-4)
    net.to(torch.float64)
    assert [id(a) for a in net.arch_parameters()] == ids           # same tensor objects: the optimiser keeps them
    assert all(a.dtype == torch.float64 and a.requires_grad for a in net.arch_parameters())
    assert opt.param_groups[0]['params'][0] is net.alphas_edges


@pytest.mark.parametrize('name', ['search_ntu_small', 'search_mmimdb_small', 'search_ego_small', 'found_mixed'])
def test_launch_plan_builds_in_validate_only_mode(name, validate_only):
    from bmnas.nn import CrossEntropyLoss, BCEWithLogitsLoss
    d = load(name)
    cfg = cfg_of(d)
    gt = unpickle_genotype(d['genotype']) if 'genotype' in d else None
    cpu = torch.device('cpu')
    head = U.build_head(cfg, int(d['num_classes']), sub(d, 'sd0/'), arch_of(d, 'arch0/') if gt is None else None,
                        genotype=gt, device=cpu)
    head.train()
    U.inject_masks(head, sub(d, 'fb/mask/'), device=cpu)
    feats = [t.requires_grad_(True) for t in arch_of(d, 'fb/feat/')]
    labels = torch.from_numpy(d['fb/labels'])
    out = head(feats)
    crit = CrossEntropyLoss() if labels.ndim == 1 else BCEWithLogitsLoss()
    crit(out, labels).backward()
    runner = list(head.fusion_net._bm_cache.values())[0]
    names_f = [c.name for c in runner.prog.fwd]
    names_b = [c.name for c in runner.prog.bwd]
    n_nodes = cfg.steps * cfg.node_steps
    assert names_f.count('bmnas_node_fwd') == n_nodes and names_b.count('bmnas_node_bwd') == n_nodes
    assert names_f[-1] == 'bmnas_ln_fwd' and names_b[0] == 'bmnas_ln_bwd'
    # every weight and architecture tensor got a gradient view inside ONE flat arena
    arena = head._bm_joint
    for p in list(head.parameters()) + head.arch_parameters():
        assert p.grad is not None and p.grad.data_ptr() == arena.view(p).data_ptr()
    lo, hi = arena.flat.data_ptr(), arena.flat.data_ptr() + arena.flat.numel() * 4
    assert all(lo <= p.grad.data_ptr() < hi for p in head.parameters())


@pytest.mark.parametrize('tag', ['ntu_5d', 'ntu_vec', 'mm_map'])
def test_reshape_layer_plan_and_state_dict(tag, validate_only):
    """drop-in ReshapeInputLayer / ReshapeInputLayer_MMIMDB: reference state_dict loads by name, the launch plan is
    pool -> conv -> node (and the mirrored backward incl. the pooling gather when the input wants a gradient)"""
    from models.auxiliary.aux_models import ReshapeInputLayer, ReshapeInputLayer_MMIMDB
    d = load('reshape')
    C, L, mm = (int(v) for v in d[f'{tag}/meta'])
    x = torch.from_numpy(d[f'{tag}/x']).requires_grad_(True)
    cls = ReshapeInputLayer_MMIMDB if mm else ReshapeInputLayer
    mod = cls(x.shape[1], C, L, types.SimpleNamespace(drpt=0.2))
    mod.load_state_dict({k[3:]: v for k, v in sub(d, f'{tag}/sd0/').items()}, strict=True)
    mod.train()
    U.inject_masks(mod, {k[3:]: v for k, v in sub(d, f'{tag}/mask/').items()}, device=torch.device('cpu'))
    out = mod(x)
    assert out.shape == (x.shape[0], C, L)
    out.backward(torch.from_numpy(d[f'{tag}/go']))
    runner = list(mod._bm_cache.values())[0]
    assert [c.name for c in runner.prog.fwd] == ['bmnas_pool_fwd', 'bmnas_conv_fwd', 'bmnas_node_fwd']
    assert [c.name for c in runner.prog.bwd] == ['bmnas_node_bwd', 'bmnas_conv_wgrad', 'bmnas_conv_dgrad', 'bmnas_pool_bwd']
    assert x.grad is not None and x.grad.shape == x.shape
    assert all(p.grad is not None for p in mod.parameters())


def test_chained_edge_mixes_remove_launches(validate_only, monkeypatch):
    """launch-plan structure: with chaining the only stand-alone forward edge mixes are the cell-level ones (the
    inner mixes are written by their producers), every inner mix keeps exactly one dot-only backward call for
    d(beta), and the main-chain input-gradient calls of the inner mixes are gone; without chaining the plan is the
    one-kernel-per-reference-op sequence"""
    from bmnas import program
    from bmnas.nn import CrossEntropyLoss
    d = load('search_ego_small')             # steps=2, node_steps=3
    cfg = cfg_of(d)
    counts = {}
    for chain in (True, False):
        monkeypatch.setattr(program, 'CHAIN_MIX', chain)
        monkeypatch.setattr(program, 'CHAIN_NODE', chain)
        head = U.build_head(cfg, int(d['num_classes']), sub(d, 'sd0/'), arch_of(d, 'arch0/'), device=torch.device('cpu'))
        head.train()
        feats = [t.requires_grad_(True) for t in arch_of(d, 'fb/feat/')]
        CrossEntropyLoss()(head(feats), torch.from_numpy(d['fb/labels'])).backward()
        prog = list(head.fusion_net._bm_cache.values())[0].prog
        f = [c.name for c in prog.fwd]
        main_b = [c.name for c in prog.bwd if not c.side]
        side_b = [c.name for c in prog.bwd if c.side]
        counts[chain] = (f.count('bmnas_mix_fwd'), main_b.count('bmnas_mix_bwd'), side_b.count('bmnas_mix_bwd'),
                         f.count('bmnas_node_fwd'))
    n_cells, ns = cfg.steps, cfg.node_steps
    assert counts[False] == (n_cells * (1 + ns), n_cells * (1 + ns), n_cells * (1 + ns), n_cells * ns)
    assert counts[True] == (n_cells, n_cells, n_cells * (1 + ns), n_cells * ns)
