"""Generate the golden fixtures in this directory by EXECUTING THE REAL REFERENCE.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

The reference modules (models/search/darts/*) are imported unmodified; the only
interventions are:
  * stub modules for the missing optional imports ``IPython`` and ``graphviz``;
  * ``nn.Dropout.forward`` is replaced by ``x * mask / (1-p)`` with recorded,
    seeded keep-masks keyed by the dropout module's qualified name (bit-matching
    torch's global Philox stream is not a goal);
  * architecture tensors are overwritten with larger random values so that the
    soft-max weights are not all ~0.5.
Everything written is a tensor the reference computed: outputs, gradients, BN
buffers, Adam-updated parameters, LR sequences, genotypes, pickled bytes.
"""
import io
import os
import zlib
import pickle
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'

ip = types.ModuleType('IPython'); ip.embed = lambda *a, **k: None; sys.modules['IPython'] = ip
gv = types.ModuleType('graphviz'); gv.Digraph = object; sys.modules['graphviz'] = gv
sys.path.insert(0, REF)
from models.search.darts.model_search import FusionNetwork          # noqa: E402  (import first, SURVEY C-11)
from models.search.darts.node_search import FusionNode, NodeCell    # noqa: E402
from models.search.darts.architect import Architect                 # noqa: E402
from models.search.darts.model import Found_FusionNetwork           # noqa: E402
from models.search.darts import node_operations as ref_nops         # noqa: E402
from models.search.darts import genotypes as ref_gt                 # noqa: E402
from models.search.darts.genotypes import Genotype, StepGenotype    # noqa: E402
import models.auxiliary.scheduler as ref_sc                         # noqa: E402
import models.auxiliary.aux_models as ref_aux                       # noqa: E402

MASKS = {}          # name -> mask, filled lazily by the patched dropout
MASK_SEED = [0]


def _patched_dropout_forward(self, x):
    if not self.training:
        return x
    name = getattr(self, '_qualname', None)
    if name is None:
        raise RuntimeError('dropout module without a registered name')
    if name not in MASKS:
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2 ** 31) + MASK_SEED[0])
        MASKS[name] = (torch.rand(x.shape, generator=g) >= self.p).to(torch.uint8)
    return x * MASKS[name].to(x.dtype) / (1.0 - self.p)


nn.Dropout.forward = _patched_dropout_forward


def name_dropouts(model, prefix=''):
    for n, m in model.named_modules():
        if isinstance(m, nn.Dropout):
            m._qualname = (prefix + n)


class Args:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.weight_decay = 3e-4
        self.parallel = False


class Head(nn.Module):
    """Searchable_*_Net minus backbones and reshape layers
    (ntu_darts_searchable.py:94-101,149-150)."""

    def __init__(self, args, num_classes, genotype=None):
        super().__init__()
        if genotype is None:
            self.fusion_net = FusionNetwork(args.steps, args.multiplier, args.num_input_nodes, 2, args,
                                            criterion=None)
            mult = args.multiplier
        else:
            self.fusion_net = Found_FusionNetwork(args.steps, args.multiplier, args.num_input_nodes, 2, args,
                                                  None, genotype)
            mult = len(genotype.concat)
        self.central_classifier = nn.Linear(args.C * args.L * mult, num_classes)

    def forward(self, feats):
        return self.central_classifier(self.fusion_net(feats))

    def arch_parameters(self):
        return self.fusion_net.arch_parameters()


def randomize(model, seed):
    """Well-conditioned random weights (BN/LN affine away from the 1/0 defaults)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('ln.weight') or n.endswith('bn.weight'):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith('ln.bias') or n.endswith('bn.bias'):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            else:
                p.add_(0.01 * torch.randn(p.shape, generator=g))


def set_arch(model, seed, scale):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for a in model.arch_parameters():
            a.copy_(scale * torch.randn(a.shape, generator=g))


def to_np(d):
    return {k: (v.detach().cpu().numpy().copy() if torch.is_tensor(v) else np.asarray(v).copy()) for k, v in d.items()}


def save(name, **arrs):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrs)
    print('wrote', path, os.path.getsize(path), 'bytes')


def feats_labels(args, B, num_classes, seed, loss):
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(B, args.C, args.L, generator=g) for _ in range(args.num_input_nodes)]
    if loss == 'ce':
        labels = torch.randint(0, num_classes, (B,), generator=g)
    else:
        labels = (torch.rand(B, num_classes, generator=g) < 0.2).float()
    return feats, labels


def crit(loss):
    return nn.CrossEntropyLoss() if loss == 'ce' else nn.BCEWithLogitsLoss()


# ---------------------------------------------------------------- search cases
def search_case(tag, B, num_classes, loss, nsteps, seed, arch_scale, **cfg):
    """fwd/bwd of the searchable head + ``nsteps`` full search steps
    (Architect.step + weight step, train_searchable/ntu.py:70-93)."""
    MASKS.clear()
    torch.manual_seed(seed)
    args = Args(**cfg)
    model = Head(args, num_classes)
    randomize(model, seed + 1)
    set_arch(model, seed + 2, arch_scale)
    name_dropouts(model)
    model.train()
    out = {'cfg_' + k: np.asarray(v) for k, v in cfg.items()}
    out['num_classes'] = np.asarray(num_classes)
    out['B'] = np.asarray(B)
    for k, v in to_np(model.state_dict()).items():
        out['sd0/' + k] = v
    for i, a in enumerate(model.arch_parameters()):
        out[f'arch0/{i}'] = a.detach().numpy().copy()

    # ---- single fwd+bwd with input grads -------------------------------------
    feats, labels = feats_labels(args, B, num_classes, seed + 3, loss)
    feats = [f.requires_grad_(True) for f in feats]
    MASK_SEED[0] = 0
    logits = model(feats)
    lv = crit(loss)(logits, labels)
    lv.backward()
    for i, f in enumerate(feats):
        out[f'fb/feat/{i}'] = f.detach().numpy()
        out[f'fb/gfeat/{i}'] = f.grad.numpy().copy()
    out['fb/labels'] = labels.numpy()
    out['fb/logits'] = logits.detach().numpy()
    out['fb/loss'] = lv.detach().numpy()
    for n, p in model.named_parameters():
        out['fb/g/' + n] = p.grad.numpy().copy()
    for i, a in enumerate(model.arch_parameters()):
        out[f'fb/ga/{i}'] = a.grad.numpy().copy()
    for k, v in MASKS.items():
        out['fb/mask/' + k] = v.numpy()
    for k, v in to_np(model.state_dict()).items():
        if 'running' in k or 'num_batches' in k:
            out['fb/sd/' + k] = v

    # ---- eval-mode forward (running statistics, no dropout) ---------------------
    model.eval()
    with torch.no_grad():
        out['eval/logits'] = model([f.detach() for f in feats]).numpy()
    model.train()

    # ---- nsteps of the search loop ---------------------------------------------
    model.zero_grad()
    for a in model.arch_parameters():
        a.grad = None
    params = [{'params': model.fusion_net.parameters()}, {'params': model.central_classifier.parameters()}]
    eta_max, eta_min, Ti, Tm, nbpe = 1e-3, 1e-6, 1, 2, 4.0
    opt = torch.optim.Adam(params, lr=eta_max, weight_decay=args.weight_decay)
    sched = ref_sc.LRCosineAnnealingScheduler(eta_max, eta_min, Ti, Tm, nbpe)
    aopt = torch.optim.Adam(model.arch_parameters(), lr=3e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    architect = Architect(model, args, crit(loss), aopt)
    out['loop/hyper'] = np.asarray([eta_max, eta_min, Ti, Tm, nbpe, args.weight_decay, 3e-4, 1e-3])
    for s in range(nsteps):
        dev = feats_labels(args, B, num_classes, seed + 10 + 2 * s, loss)
        trn = feats_labels(args, B, num_classes, seed + 11 + 2 * s, loss)
        MASKS.clear(); MASK_SEED[0] = 100 + 2 * s
        architect.step(dev[0], dev[1], None)
        for k, v in MASKS.items():
            out[f'loop/{s}/mask_dev/' + k] = v.numpy()
        opt.zero_grad()
        MASKS.clear(); MASK_SEED[0] = 101 + 2 * s
        logits = model(trn[0])
        lv = crit(loss)(logits, trn[1])
        sched.step(); sched.update_optimizer(opt)
        lv.backward()
        opt.step()
        for k, v in MASKS.items():
            out[f'loop/{s}/mask_train/' + k] = v.numpy()
        for i in range(len(dev[0])):
            out[f'loop/{s}/dev_feat/{i}'] = dev[0][i].numpy()
            out[f'loop/{s}/train_feat/{i}'] = trn[0][i].numpy()
        out[f'loop/{s}/dev_labels'] = dev[1].numpy()
        out[f'loop/{s}/train_labels'] = trn[1].numpy()
        out[f'loop/{s}/train_loss'] = lv.detach().numpy()
        out[f'loop/{s}/lr'] = np.asarray(sched.eta)
        out[f'loop/{s}/genotype'] = np.frombuffer(pickle.dumps(model.fusion_net.genotype()), dtype=np.uint8)
        for i, a in enumerate(model.arch_parameters()):
            out[f'loop/{s}/arch/{i}'] = a.detach().numpy().copy()
    for k, v in to_np(model.state_dict()).items():
        out['loop/sd_final/' + k] = v
    out['nsteps'] = np.asarray(nsteps)
    save(tag, **out)


# ---------------------------------------------------------------- found cases
def found_case(tag, B, num_classes, genotype, seed, **cfg):
    MASKS.clear(); MASK_SEED[0] = 7
    torch.manual_seed(seed)
    args = Args(**cfg)
    model = Head(args, num_classes, genotype)
    randomize(model, seed + 1)
    name_dropouts(model)
    model.train()
    out = {'cfg_' + k: np.asarray(v) for k, v in cfg.items()}
    out['num_classes'] = np.asarray(num_classes)
    out['genotype'] = np.frombuffer(pickle.dumps(genotype), dtype=np.uint8)
    for k, v in to_np(model.state_dict()).items():
        out['sd0/' + k] = v
    feats, labels = feats_labels(args, B, num_classes, seed + 3, 'ce')
    feats = [f.requires_grad_(True) for f in feats]
    logits = model(feats)
    lv = nn.CrossEntropyLoss()(logits, labels)
    lv.backward()
    for i, f in enumerate(feats):
        out[f'fb/feat/{i}'] = f.detach().numpy()
        out[f'fb/gfeat/{i}'] = (f.grad.numpy().copy() if f.grad is not None else np.zeros(f.shape, np.float32))
    out['fb/labels'] = labels.numpy()
    out['fb/logits'] = logits.detach().numpy()
    out['fb/loss'] = lv.detach().numpy()
    for n, p in model.named_parameters():
        out['fb/g/' + n] = p.grad.numpy().copy()
    for k, v in MASKS.items():
        out['fb/mask/' + k] = v.numpy()
    for k, v in to_np(model.state_dict()).items():
        if 'running' in k or 'num_batches' in k:
            out['fb/sd/' + k] = v
    model.eval()
    with torch.no_grad():
        out['eval/logits'] = model([f.detach() for f in feats]).numpy()
    save(tag, **out)


# ---------------------------------------------------------------- reshape layers
RESHAPE_CASES = [
    # tag, class, raw feature shape, C, L          (shapes follow SURVEY App. B, channel counts scaled down)
    ('ntu_5d', 'ReshapeInputLayer', (3, 12, 8, 3, 3), 16, 8),        # (B, C_in, T=L, H, W): pool the spatial dims only
    ('ntu_skel', 'ReshapeInputLayer', (3, 10, 4, 4), 16, 8),         # d2 = 4 < L: adaptive bins repeat rows (up-sampling)
    ('ntu_vec', 'ReshapeInputLayer', (3, 20), 16, 8),                # pooled (B, C_in) vector: replicated over L
    ('ragged_down', 'ReshapeInputLayer', (4, 6, 11, 5), 8, 8),       # d2 = 11 -> 8 overlapping bins
    ('ragged_up', 'ReshapeInputLayer', (4, 6, 5, 7), 8, 8),          # d2 = 5 -> 8
    ('ego_5d', 'ReshapeInputLayer', (2, 9, 2, 4, 4), 16, 8),
    ('mm_map', 'ReshapeInputLayer_MMIMDB', (3, 8, 7, 7), 24, 16),    # VGG feature map -> 4 x 4
    ('mm_vec', 'ReshapeInputLayer_MMIMDB', (3, 9), 24, 16),          # Maxout vector -> replicated
    ('mm_small', 'ReshapeInputLayer_MMIMDB', (3, 5, 3, 2), 12, 4),   # 3 x 2 -> 2 x 2
]


def reshape_cases():
    """ReshapeInputLayer / ReshapeInputLayer_MMIMDB (models/auxiliary/aux_models.py:51-115): forward, all gradients
    (input included), BatchNorm buffers after the step, eval-mode forward."""
    out = {}
    for ci, (tag, cls, shape, C, L) in enumerate(RESHAPE_CASES):
        MASKS.clear()
        MASK_SEED[0] = 900 + ci
        g = torch.Generator().manual_seed(300 + ci)
        args = Args(drpt=0.2)
        mod = getattr(ref_aux, cls)(shape[1], C, L, args)
        randomize(mod, 310 + ci)
        name_dropouts(mod, prefix='op.')
        mod.train()
        for k, v in to_np(mod.state_dict()).items():
            out[f'{tag}/sd0/op.{k}'] = v
        x = torch.randn(*shape, generator=g).requires_grad_(True)
        y = mod(x)
        go = torch.randn(y.shape, generator=g)
        y.backward(go)
        out[f'{tag}/x'] = x.detach().numpy()
        out[f'{tag}/out'] = y.detach().numpy()
        out[f'{tag}/go'] = go.numpy()
        out[f'{tag}/gx'] = x.grad.numpy().copy()
        for n, p_ in mod.named_parameters():
            out[f'{tag}/g/op.{n}'] = p_.grad.numpy().copy()
        for k, v in MASKS.items():
            out[f'{tag}/mask/{k}'] = v.numpy()
        for k, v in to_np(mod.state_dict()).items():
            if 'running' in k or 'num_batches' in k:
                out[f'{tag}/sd1/op.{k}'] = v
        mod.eval()
        with torch.no_grad():
            out[f'{tag}/eval_out'] = mod(x.detach()).numpy()
        out[f'{tag}/meta'] = np.asarray([C, L, int(cls.endswith('MMIMDB'))])
    save('reshape', **out)


# ---------------------------------------------------------------- primitives
def primitive_cases():
    """Each step-node primitive alone, x != y, train + eval, incl. the unregistered
    CatConvMish (node_operations.py:66-82) and a 5-op NodeMixedOp with it appended."""
    C, L, B = 16, 8, 5
    args = Args(C=C, L=L, drpt=0.2)
    out = {}
    g = torch.Generator().manual_seed(5)
    mods = {
        'Sum': ref_nops.Sum(), 'ScaleDotAttn': ref_nops.ScaledDotAttn(C, L),
        'LinearGLU': ref_nops.LinearGLU(C, args), 'ConcatFC': ref_nops.ConcatFC(C, args),
        'CatConvMish': ref_nops.CatConvMish(C, args),
    }
    for name, m in mods.items():
        MASKS.clear(); MASK_SEED[0] = 11
        randomize(m, 3)
        name_dropouts(m, prefix='op.')
        for dn, dm in m.named_modules():
            if isinstance(dm, nn.Dropout):
                dm._qualname = 'op.' + dn
        m.train()
        x = torch.randn(B, C, L, generator=g, requires_grad=True)
        y = torch.randn(B, C, L, generator=g, requires_grad=True)
        go = torch.randn(B, C, L, generator=g)
        for k, v in to_np(m.state_dict()).items():
            out[f'{name}/sd0/op.{k}'] = v
        o = m(x, y)
        o.backward(go)
        out[f'{name}/x'], out[f'{name}/y'], out[f'{name}/go'] = x.detach().numpy(), y.detach().numpy(), go.numpy()
        out[f'{name}/out'] = o.detach().numpy()
        out[f'{name}/gx'], out[f'{name}/gy'] = x.grad.numpy(), y.grad.numpy()
        for n, p in m.named_parameters():
            out[f'{name}/g/op.{n}'] = p.grad.numpy().copy()
        for k, v in MASKS.items():
            out[f'{name}/mask/{k}'] = v.numpy()
        for k, v in to_np(m.state_dict()).items():
            out[f'{name}/sd1/op.{k}'] = v
        m.eval()
        with torch.no_grad():
            out[f'{name}/eval_out'] = m(x.detach(), y.detach()).numpy()

    # 5-op mixed op: register CatConvMish at run time
    ref_nops.STEP_STEP_OPS['CatConvMish'] = lambda C, L, args: ref_nops.CatConvMish(C, args)
    ref_nops.STEP_STEP_PRIMITIVES.append('CatConvMish')
    try:
        MASKS.clear(); MASK_SEED[0] = 13
        m = ref_nops.NodeMixedOp(C, L, args)
        randomize(m, 4)
        name_dropouts(m, prefix='mix.')
        m.train()
        x = torch.randn(B, C, L, generator=g, requires_grad=True)
        y = torch.randn(B, C, L, generator=g, requires_grad=True)
        w = torch.softmax(torch.randn(5, generator=g), -1).requires_grad_(True)
        go = torch.randn(B, C, L, generator=g)
        for k, v in to_np(m.state_dict()).items():
            out[f'Mixed5/sd0/mix.{k}'] = v
        o = m(x, y, w)
        o.backward(go)
        out['Mixed5/x'], out['Mixed5/y'], out['Mixed5/go'] = x.detach().numpy(), y.detach().numpy(), go.numpy()
        out['Mixed5/w'], out['Mixed5/gw'] = w.detach().numpy(), w.grad.numpy()
        out['Mixed5/out'] = o.detach().numpy()
        out['Mixed5/gx'], out['Mixed5/gy'] = x.grad.numpy(), y.grad.numpy()
        for n, p in m.named_parameters():
            out[f'Mixed5/g/mix.{n}'] = p.grad.numpy().copy()
        for k, v in MASKS.items():
            out[f'Mixed5/mask/{k}'] = v.numpy()
    finally:
        ref_nops.STEP_STEP_PRIMITIVES.pop()
        del ref_nops.STEP_STEP_OPS['CatConvMish']
    save('primitives', **out)


# ---------------------------------------------------------------- genotype / schedule / pickle
def genotype_cases():
    out = {}
    idx = 0
    for steps, mult, n_in, ns, nm, scale, seed in [
        (2, 2, 8, 2, 2, 1.0, 1), (2, 2, 8, 2, 2, 1e-3, 2), (3, 3, 8, 3, 3, 1.0, 3), (4, 4, 8, 3, 3, 0.5, 4),
        (2, 2, 6, 1, 1, 1.0, 5), (2, 2, 8, 2, 2, 0.0, 6), (2, 2, 4, 2, 1, 1.0, 7), (2, 2, 8, 3, 3, 2.0, 8),
    ]:
        args = Args(C=8, L=4, drpt=0.1, num_input_nodes=n_in, node_steps=ns, node_multiplier=nm)
        net = FusionNetwork(steps, mult, n_in, 2, args)
        set_arch(net, seed, scale)
        if seed == 7:   # exact ties between the two best input nodes
            with torch.no_grad():
                net.alphas_edges[:, :] = 0.0
                net.alphas_edges[1, 1] = 1.0
                net.alphas_edges[2, 1] = 1.0
                net.alphas_edges[3, 1] = 1.0
        out[f'{idx}/cfg'] = np.asarray([steps, mult, n_in, ns, nm])
        for i, a in enumerate(net.arch_parameters()):
            out[f'{idx}/arch/{i}'] = a.detach().numpy().copy()
        gt = net.genotype()
        out[f'{idx}/pickle'] = np.frombuffer(pickle.dumps(gt), dtype=np.uint8)
        out[f'{idx}/str'] = np.asarray(str(gt))
        idx += 1
    out['n'] = np.asarray(idx)
    save('genotypes', **out)


def scheduler_cases():
    out = {}
    for i, (emax, emin, Ti, Tm, nbpe) in enumerate([(1e-3, 1e-6, 1, 2, 7.5), (3e-3, 1e-6, 5, 2, 3.0),
                                                    (1e-3, 1e-6, 1, 2, 420.3)]):
        sc = ref_sc.LRCosineAnnealingScheduler(emax, emin, Ti, Tm, nbpe)
        out[f'{i}/hyper'] = np.asarray([emax, emin, Ti, Tm, nbpe])
        out[f'{i}/lr'] = np.asarray([sc.step() for _ in range(400)])
    out['n'] = np.asarray(3)
    save('scheduler', **out)


def adam_cases():
    """torch.optim.Adam with both hyper-parameter sets the path uses."""
    out = {}
    g = torch.Generator().manual_seed(9)
    for tag, (lr, betas, wd) in {'weight': (1e-3, (0.9, 0.999), 3e-4), 'arch': (3e-4, (0.5, 0.999), 1e-3)}.items():
        ps = [torch.randn(7, 5, generator=g).requires_grad_(True), torch.randn(33, generator=g).requires_grad_(True)]
        opt = torch.optim.Adam(ps, lr=lr, betas=betas, weight_decay=wd)
        out[f'{tag}/hyper'] = np.asarray([lr, betas[0], betas[1], wd])
        out[f'{tag}/p0/0'], out[f'{tag}/p0/1'] = ps[0].detach().numpy().copy(), ps[1].detach().numpy().copy()
        for s in range(20):
            for j, p in enumerate(ps):
                p.grad = torch.randn(p.shape, generator=g) * (0.1 if s % 3 else 1e-6)
                out[f'{tag}/g/{s}/{j}'] = p.grad.numpy().copy()
            opt.step()
            for j, p in enumerate(ps):
                out[f'{tag}/p/{s}/{j}'] = p.detach().numpy().copy()
    save('adam', **out)


NTU_GOLDEN = Genotype(edges=[('skip', 2), ('skip', 7), ('skip', 2), ('skip', 3)],
                      steps=[StepGenotype(inner_edges=[('skip', 0), ('skip', 1), ('skip', 2), ('skip', 0)],
                                          inner_steps=['LinearGLU', 'LinearGLU'], inner_concat=[2, 3]),
                             StepGenotype(inner_edges=[('skip', 0), ('skip', 1), ('skip', 2), ('skip', 0)],
                                          inner_steps=['ScaleDotAttn', 'ScaleDotAttn'], inner_concat=[2, 3])],
                      concat=[8, 9])   # visualize.ipynb:618, structure_vis.ipynb:145
MIXED_FOUND = Genotype(edges=[('skip', 1), ('skip', 3), ('skip', 0), ('skip', 4)],
                       steps=[StepGenotype(inner_edges=[('skip', 1), ('skip', 0), ('skip', 2), ('skip', 1)],
                                           inner_steps=['ConcatFC', 'ScaleDotAttn'], inner_concat=[2, 3]),
                              StepGenotype(inner_edges=[('skip', 0), ('skip', 1), ('none', 2), ('skip', 0)],
                                           inner_steps=['Sum', 'LinearGLU'], inner_concat=[2, 3])],
                       concat=[4, 5])
FOUND_NM1 = Genotype(edges=[('skip', 2), ('skip', 0), ('skip', 1), ('skip', 3)],
                     steps=[StepGenotype(inner_edges=[('skip', 1), ('skip', 0)], inner_steps=['ConcatFC'], inner_concat=[2]),
                            StepGenotype(inner_edges=[('skip', 0), ('skip', 1)], inner_steps=['LinearGLU'], inner_concat=[2])],
                     concat=[3, 4])    # shape of visualize.ipynb:352 (MM-IMDB, old op names mapped)

if __name__ == '__main__':
    torch.set_num_threads(4)
    search_case('search_ntu_small', B=6, num_classes=7, loss='ce', nsteps=3, seed=20, arch_scale=0.7,
                C=16, L=8, num_input_nodes=4, steps=2, multiplier=2, node_steps=2, node_multiplier=2, drpt=0.2)
    search_case('search_mmimdb_small', B=5, num_classes=6, loss='bce', nsteps=2, seed=30, arch_scale=0.5,
                C=24, L=16, num_input_nodes=3, steps=2, multiplier=2, node_steps=1, node_multiplier=1, drpt=0.1)
    search_case('search_ego_small', B=4, num_classes=5, loss='ce', nsteps=2, seed=40, arch_scale=1e-3,
                C=16, L=8, num_input_nodes=4, steps=2, multiplier=2, node_steps=3, node_multiplier=3, drpt=0.05)
    search_case('search_deep_small', B=4, num_classes=5, loss='ce', nsteps=1, seed=50, arch_scale=0.3,
                C=32, L=4, num_input_nodes=6, steps=3, multiplier=3, node_steps=2, node_multiplier=2, drpt=0.2)
    found_case('found_ntu_golden', B=6, num_classes=7, genotype=NTU_GOLDEN, seed=60,
               C=16, L=8, num_input_nodes=8, steps=2, multiplier=2, node_steps=2, node_multiplier=2, drpt=0.2)
    found_case('found_mixed', B=5, num_classes=4, genotype=MIXED_FOUND, seed=70,
               C=16, L=8, num_input_nodes=4, steps=2, multiplier=2, node_steps=2, node_multiplier=2, drpt=0.2)
    found_case('found_nm1', B=5, num_classes=4, genotype=FOUND_NM1, seed=80,
               C=24, L=16, num_input_nodes=4, steps=2, multiplier=2, node_steps=1, node_multiplier=1, drpt=0.1)
    primitive_cases()
    reshape_cases()
    genotype_cases()
    scheduler_cases()
    adam_cases()
