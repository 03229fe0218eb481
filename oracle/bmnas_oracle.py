"""CPU oracle for the BM-NAS search-step hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional (module-free) restatement, in plain PyTorch-on-CPU
tensor arithmetic, of the algorithm the reference implements with nn.Modules
under ``models/search/darts/``.  It is the checker for the CUDA path; nothing
under ``bm-nas_b200/`` may import it.  Only ``tests/``, ``__graft_entry__.smoke``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it.

Parity pinning: ``tests/golden/*.npz`` were produced by importing the real
reference modules from /root/reference (script: ``tests/golden/make_golden.py``)
and ``tests/test_oracle_golden.py`` checks every function here against them
(outputs, all gradients, BatchNorm running statistics, Adam updates, LR
schedule, genotype derivation, pickled genotype bytes).

Conventions
-----------
* ``P``      dict name -> tensor, keyed exactly like the reference ``state_dict``
             (e.g. ``cell._step_nodes.0.node_cell.node_ops.1._ops.2.conv.weight``).
             BatchNorm buffers in ``P`` are updated in place in training mode.
* ``masks``  dict dropout-module-path -> keep mask (0/1, same shape as the
             dropout input).  ``None``/missing in training mode means "no
             dropout at all" (identity), a mask means ``x * mask / (1-p)``.
* tensors are (B, C, L); dtype follows the inputs (fp32 or the fp64 referee).

All citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
import pickle
from collections import namedtuple
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# genotype format -- models/search/darts/genotypes.py:3-21
Genotype = namedtuple('Genotype', 'edges steps concat')
StepGenotype = namedtuple('StepGenotype', 'inner_edges inner_steps inner_concat')
PRIMITIVES = ['none', 'skip']
STEP_EDGE_PRIMITIVES = ['none', 'skip']
STEP_STEP_PRIMITIVES = ['Sum', 'ScaleDotAttn', 'LinearGLU', 'ConcatFC']

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LN_EPS = 1e-5
ATTN_DROP = 0.1  # node_operations.py:89 (hard-coded)


class Cfg:
    """The ``args`` fields the path consumes."""

    def __init__(self, C, L, num_input_nodes, steps, multiplier, node_steps,
                 node_multiplier, drpt, step_ops: Sequence[str] = STEP_STEP_PRIMITIVES):
        self.C, self.L = C, L
        self.num_input_nodes = num_input_nodes
        self.steps, self.multiplier = steps, multiplier
        self.node_steps, self.node_multiplier = node_steps, node_multiplier
        self.drpt = drpt
        self.step_ops = list(step_ops)


# --------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------
def _dropout(x, p, masks, name, training):
    """nn.Dropout with an injected keep-mask (SURVEY 8c harness rule i)."""
    if not training or masks is None:
        return x
    m = masks.get(name)
    if m is None:
        return x
    return x * m.to(x.dtype) / (1.0 - p)


def batchnorm(z, P, prefix, training):
    """nn.BatchNorm1d over (B, C, L): batch statistics in training mode, running
    statistics otherwise; running_var gets the unbiased variance
    (node_operations.py:27,34)."""
    w, b = P[prefix + '.weight'], P[prefix + '.bias']
    if training:
        n = z.shape[0] * z.shape[2]
        mean = z.mean(dim=(0, 2))
        var = ((z - mean[None, :, None]) ** 2).mean(dim=(0, 2))
        with torch.no_grad():
            rm, rv = P[prefix + '.running_mean'], P[prefix + '.running_var']
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().to(rm.dtype))
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (var.detach() * n / max(n - 1, 1)).to(rv.dtype))
            P[prefix + '.num_batches_tracked'].add_(1)
    else:
        mean = P[prefix + '.running_mean'].to(z.dtype)
        var = P[prefix + '.running_var'].to(z.dtype)
    zh = (z - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS)
    return zh * w[None, :, None] + b[None, :, None]


def layernorm_cl(z, w, b):
    """nn.LayerNorm([C', L]) -- per sample over all C'*L elements, affine (C', L)."""
    mean = z.mean(dim=(1, 2), keepdim=True)
    var = ((z - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    return (z - mean) / torch.sqrt(var + LN_EPS) * w[None] + b[None]


def conv1x1(u, w, b):
    """nn.Conv1d(k=1): z[b] = W u[b] + bias.  w is (Cout, Cin, 1)."""
    return torch.einsum('mk,bkl->bml', w[:, :, 0], u) + b[None, :, None]


def adaptive_max_pool_2d(x4, out_h, out_w):
    """nn.AdaptiveMaxPool2d over the last two dims of (B, C, H, W): output bin (i, j) covers rows
    [floor(i*H/out_h), ceil((i+1)*H/out_h)) and columns likewise (ATen adaptive pooling start/end index rule);
    bins overlap or repeat when H is not a multiple of out_h (also used to UP-sample: H < out_h)."""
    B, C, H, W = x4.shape
    rows = []
    for i in range(out_h):
        h0, h1 = (i * H) // out_h, -((-(i + 1) * H) // out_h)
        cols = []
        for j in range(out_w):
            w0, w1 = (j * W) // out_w, -((-(j + 1) * W) // out_w)
            cols.append(x4[:, :, h0:h1, w0:w1].amax(dim=(2, 3)))
        rows.append(torch.stack(cols, dim=-1))
    return torch.stack(rows, dim=-2)             # (B, C, out_h, out_w)


def reshape_input(x, P, prefix, L, masks, training, drpt, mmimdb=False):
    """ReshapeInputLayer.forward (models/auxiliary/aux_models.py:61-76) and ReshapeInputLayer_MMIMDB.forward
    (:102-115): raw backbone feature (B, C_in[, d2[, ...]]) -> (B, C, L).
      NTU/Ego: view (B, C_in, d2, rest) -> AdaptiveMaxPool2d((L, 1)) -> F.interpolate(size=L) (nearest, same
               length: the identity) -> Conv1d(C_in, C, 1) -> BatchNorm1d -> ReLU -> Dropout(drpt)
      MM-IMDB: view (B, C_in, d2, rest) -> AdaptiveMaxPool2d((sqrt L, sqrt L)) -> flatten -> the same conv block."""
    B, Cin = x.shape[0], x.shape[1]
    if mmimdb:
        x4 = x.reshape(B, Cin, 1, 1) if x.dim() == 2 else x.reshape(B, Cin, x.shape[2], -1)
        ps = int(math.sqrt(L))
        assert ps * ps == L
        pooled = adaptive_max_pool_2d(x4, ps, ps).reshape(B, Cin, L)
    else:
        x4 = x.reshape(B, Cin, 1, 1) if x.dim() == 2 else x.reshape(B, Cin, x.shape[2], -1)
        pooled = adaptive_max_pool_2d(x4, L, 1).reshape(B, Cin, L)
    z = conv1x1(pooled, P[prefix + '.conv.weight'], P[prefix + '.conv.bias'])
    z = batchnorm(z, P, prefix + '.bn', training)
    z = F.relu(z)
    return _dropout(z, drpt, masks, prefix + '.dropout', training)


def edge_mix(states, w):
    """Sum of FusionMixedOps over candidate input tensors, PRIMITIVES = [none, skip]
    (operations.py:104-106, Zero :18-20, Identity :92-93; python ``sum`` order
    model_search.py:58).  ``w`` is (n, 2), already soft-maxed."""
    out = 0
    for j, h in enumerate(states):
        e = 0
        e = e + w[j][0] * h.mul(0.)
        e = e + w[j][1] * h
        out = out + e
    return out


# --------------------------------------------------------------------------
# step-node primitives -- node_operations.py
# --------------------------------------------------------------------------
def op_sum(x, y):
    """Sum.forward node_operations.py:19-20."""
    return x + y


def op_attn(x, y, P, prefix, masks, training):
    """ScaledDotAttn.forward node_operations.py:92-108."""
    q = x.transpose(1, 2)            # (B, L, C)
    k = y                            # (B, C, L)
    v = y.transpose(1, 2)            # (B, L, C)
    scores = torch.matmul(q, k) / math.sqrt(q.size(-1))
    attn = F.softmax(scores, dim=-1)
    out = torch.matmul(attn, v).transpose(1, 2)
    out = _dropout(out, ATTN_DROP, masks, prefix + '.dropout', training)
    return layernorm_cl(out, P[prefix + '.ln.weight'], P[prefix + '.ln.bias'])


def op_linear_glu(x, y, P, prefix, masks, training, drpt):
    """LinearGLU.forward node_operations.py:30-39."""
    z = conv1x1(torch.cat([x, y], dim=1), P[prefix + '.conv.weight'], P[prefix + '.conv.bias'])
    z = batchnorm(z, P, prefix + '.bn', training)
    C = z.shape[1] // 2
    out = z[:, :C] * torch.sigmoid(z[:, C:])
    return _dropout(out, drpt, masks, prefix + '.dropout', training)


def _mish(z):
    return z * torch.tanh(F.softplus(z))


def op_concat_fc(x, y, P, prefix, masks, training, drpt, act='relu'):
    """ConcatFC.forward node_operations.py:49-56; act='mish' gives CatConvMish :75-82."""
    z = conv1x1(torch.cat([x, y], dim=1), P[prefix + '.conv.weight'], P[prefix + '.conv.bias'])
    z = batchnorm(z, P, prefix + '.bn', training)
    z = F.relu(z) if act == 'relu' else _mish(z)
    return _dropout(z, drpt, masks, prefix + '.dropout', training)


def step_op(name, x, y, P, prefix, masks, training, drpt):
    if name == 'Sum':
        return op_sum(x, y)
    if name == 'ScaleDotAttn':
        return op_attn(x, y, P, prefix, masks, training)
    if name == 'LinearGLU':
        return op_linear_glu(x, y, P, prefix, masks, training, drpt)
    if name == 'ConcatFC':
        return op_concat_fc(x, y, P, prefix, masks, training, drpt, 'relu')
    if name == 'CatConvMish':
        return op_concat_fc(x, y, P, prefix, masks, training, drpt, 'mish')
    raise KeyError(name)


def node_mixed(x, y, gw, P, prefix, masks, training, cfg):
    """NodeMixedOp.forward node_operations.py:118-120: python-sum of w_k * op_k(x, y)."""
    out = 0
    for k, name in enumerate(cfg.step_ops):
        out = out + gw[k] * step_op(name, x, y, P, f'{prefix}._ops.{k}', masks, training, cfg.drpt)
    return out


# --------------------------------------------------------------------------
# NodeCell / FusionNode / FusionCell / FusionNetwork (search mode)
# --------------------------------------------------------------------------
def _node_tail(states, x, P, prefix, masks, training, cfg):
    """cat -> [conv1x1 -> BN -> ReLU -> dropout] -> += x -> LayerNorm
    (node_search.py:59-68, node.py:65-74)."""
    nm = cfg.node_multiplier
    out = torch.cat(states[-nm:], dim=1)
    if nm != 1:
        out = conv1x1(out, P[prefix + '.out_conv.weight'], P[prefix + '.out_conv.bias'])
        out = batchnorm(out, P, prefix + '.bn', training)
        out = F.relu(out)
        out = _dropout(out, cfg.drpt, masks, prefix + '.out_dropout', training)
    out = out + x
    return layernorm_cl(out, P[prefix + '.ln.weight'], P[prefix + '.ln.bias'])


def node_cell(x, y, edge_w, node_w, P, prefix, masks, training, cfg):
    """NodeCell.forward node_search.py:48-70.  edge_w (k,2) and node_w (ns,n_ops)
    are soft-maxed weights; ``prefix`` ends in ``.node_cell``."""
    states = [x, y]
    offset = 0
    for i in range(cfg.node_steps):
        t = edge_mix(states, edge_w[offset:offset + len(states)])
        s = node_mixed(t, t, node_w[i], P, f'{prefix}.node_ops.{i}', masks, training, cfg)
        offset += len(states)
        states.append(s)
    return _node_tail(states, x, P, prefix, masks, training, cfg)


def fusion_node(x, y, betas, gammas, P, prefix, masks, training, cfg):
    """FusionNode.forward node_search.py:101-105 (prefix ends in ``_step_nodes.i``)."""
    return node_cell(x, y, F.softmax(betas, dim=-1), F.softmax(gammas, dim=-1),
                     P, prefix + '.node_cell', masks, training, cfg)


def _cell_tail(states, P, prefix, cfg):
    """cat -> LayerNorm([C*m, L]) -> ReLU -> flatten (model_search.py:63-67)."""
    out = torch.cat(states[-cfg.multiplier:], dim=1)
    out = layernorm_cl(out, P[prefix + '.ln.weight'], P[prefix + '.ln.bias'])
    out = F.relu(out)
    return out.reshape(out.size(0), -1)


def fusion_network(feats, arch, P, masks, training, cfg, prefix='cell'):
    """FusionNetwork.forward + FusionCell.forward (model_search.py:93-97, 50-68).
    ``arch`` = [alphas, betas_0, gammas_0, betas_1, gammas_1, ...] (the order of
    FusionNetwork.arch_parameters(), model_search.py:90)."""
    assert len(feats) == cfg.num_input_nodes
    weights = F.softmax(arch[0], dim=-1)
    states = list(feats)
    offset = 0
    for i in range(cfg.steps):
        s_in = edge_mix(states, weights[offset:offset + len(states)])
        s = fusion_node(s_in, s_in, arch[1 + 2 * i], arch[2 + 2 * i], P,
                        f'{prefix}._step_nodes.{i}', masks, training, cfg)
        offset += len(states)
        states.append(s)
    return _cell_tail(states, P, prefix, cfg)


# --------------------------------------------------------------------------
# found (fixed genotype) network -- model.py:133-160, node.py:45-76
# --------------------------------------------------------------------------
def _edge_fixed(name, h):
    return h if name == 'skip' else h.mul(0.)


def found_node_cell(x, y, step_gene, P, prefix, masks, training, cfg):
    states = [x, y]
    for i in range(cfg.node_steps):
        (n0, i0), (n1, i1) = step_gene.inner_edges[2 * i], step_gene.inner_edges[2 * i + 1]
        a = _edge_fixed(n0, states[i0])
        b = _edge_fixed(n1, states[i1])
        s = step_op(step_gene.inner_steps[i], a, b, P, f'{prefix}.node_ops.{i}', masks, training, cfg.drpt)
        states.append(s)
    return _node_tail(states, x, P, prefix, masks, training, cfg)


def found_network(feats, genotype, P, masks, training, cfg, prefix='cell'):
    states = list(feats)
    steps = len(genotype.edges) // 2
    for i in range(steps):
        (n0, i0), (n1, i1) = genotype.edges[2 * i], genotype.edges[2 * i + 1]
        h1 = _edge_fixed(n0, states[i0])
        h2 = _edge_fixed(n1, states[i1])
        s = found_node_cell(h1, h2, genotype.steps[i], P, f'{prefix}._step_nodes.{i}.node_cell',
                            masks, training, cfg)
        states.append(s)

    class _C:  # multiplier comes from the genotype (model.py:112)
        multiplier = len(genotype.concat)
    return _cell_tail(states, P, prefix, _C)


# --------------------------------------------------------------------------
# head, loss, optimiser, schedule
# --------------------------------------------------------------------------
def head_logits(feats, arch, P, masks, training, cfg, genotype=None):
    """Searchable_*_Net.forward minus backbones/reshape layers
    (ntu_darts_searchable.py:149-150): fusion net then central_classifier."""
    if genotype is None:
        h = fusion_network(feats, arch, P, masks, training, cfg, prefix='fusion_net.cell')
    else:
        h = found_network(feats, genotype, P, masks, training, cfg, prefix='fusion_net.cell')
    return h @ P['central_classifier.weight'].t() + P['central_classifier.bias']


def adam_step(params, grads, state, lr, betas, weight_decay, eps=1e-8):
    """torch.optim.Adam (coupled L2 decay, bias correction), SURVEY App. A /
    ntu_darts_searchable.py:42,46-47.  ``state`` is a dict holding 'step' and
    per-tensor 'm'/'v' lists; updated in place, as are ``params``."""
    b1, b2 = betas
    if 'step' not in state:
        state['step'] = 0
        state['m'] = [torch.zeros_like(p) for p in params]
        state['v'] = [torch.zeros_like(p) for p in params]
    state['step'] += 1
    t = state['step']
    bc1 = 1 - b1 ** t
    bc2 = 1 - b2 ** t
    with torch.no_grad():
        for p, g, m, v in zip(params, grads, state['m'], state['v']):
            if g is None:
                continue
            g = g + weight_decay * p if weight_decay != 0 else g
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(m, denom, value=-(lr / bc1))


class CosineRestartLR:
    """LRCosineAnnealingScheduler models/auxiliary/scheduler.py:12-40."""

    def __init__(self, eta_max, eta_min, Ti, Tm, nbpe):
        self.eta_max, self.eta_min, self.Ti, self.Tm, self.nbpe = eta_max, eta_min, Ti, Tm, nbpe
        self.Tcur, self.it, self.eta = 0.0, 0.0, eta_max

    def step(self):
        self.Tcur = self.it / self.nbpe
        self.it += 1.0
        self.eta = self.eta_min + 0.5 * (self.eta_max - self.eta_min) * (1 + np.cos(np.pi * self.Tcur / self.Ti))
        eta = self.eta
        if eta <= self.eta_min + 1e-10:
            self.Tcur, self.Ti, self.it = 0, self.Ti * self.Tm, 0
        return eta


# --------------------------------------------------------------------------
# genotype derivation -- model_search.py:111-181, node_search.py:110-163
# --------------------------------------------------------------------------
def node_genotype(betas, gammas, cfg):
    ew = F.softmax(betas.detach().float().cpu(), dim=-1)
    nw = F.softmax(gammas.detach().float().cpu(), dim=-1)
    none = STEP_EDGE_PRIMITIVES.index('none')
    edge_gene, node_gene = [], []
    start, n = 0, 2
    for i in range(cfg.node_steps):
        W = ew[start:start + n]
        # stable sort on -max(non-none weight); torch scalars compare like floats
        order = sorted(range(i + 2), key=lambda r: -max(W[r][k] for k in range(len(W[r])) if k != none))[:2]
        for j in order:
            k_best = None
            for k in range(len(W[j])):
                if k != none and (k_best is None or W[j][k] > W[j][k_best]):
                    k_best = k
            edge_gene.append((STEP_EDGE_PRIMITIVES[k_best], j))
        start += n
        n += 1
    for i in range(cfg.node_steps):
        W = nw[i]
        k_best = None
        for k in range(len(W)):
            if k_best is None or W[k] > W[k_best]:
                k_best = k
        node_gene.append(cfg.step_ops[k_best])
    concat = list(range(2 + cfg.node_steps - cfg.node_multiplier, cfg.node_steps + 2))
    return StepGenotype(inner_edges=edge_gene, inner_steps=node_gene, inner_concat=concat)


def network_genotype(arch, cfg):
    """'sample strategy v3' (model_search.py:128-156): per step choose the best
    pair of ORIGINAL input nodes with at least one not yet used."""
    weights = F.softmax(arch[0].detach().float().cpu(), dim=-1).numpy()
    none = PRIMITIVES.index('none')
    n_in = cfg.num_input_nodes
    gene, selected = [], []
    start, n = 0, n_in
    for i in range(cfg.steps):
        W = weights[start:start + n].copy()
        pairs = []
        for j in range(n_in):
            for k in range(j + 1, n_in):
                if (j not in selected) or (k not in selected):
                    wj = max(W[j][t] for t in range(len(W[j])) if t != none)
                    wk = max(W[k][t] for t in range(len(W[k])) if t != none)
                    pairs.append([j, k, wj * wk])
        best = sorted(pairs, key=lambda p: -p[2])[0]
        edges = best[0:2]
        selected = list(set(selected + edges))
        for j in edges:
            k_best = None
            for k in range(len(W[j])):
                if k != none and (k_best is None or W[j][k] > W[j][k_best]):
                    k_best = k
            gene.append((PRIMITIVES[k_best], j))
        start += n
        n += 1
    steps = [node_genotype(arch[1 + 2 * i], arch[2 + 2 * i], cfg) for i in range(cfg.steps)]
    concat = list(range(n_in + cfg.steps - cfg.multiplier, cfg.steps + n_in))
    return Genotype(edges=gene, steps=steps, concat=concat)


# --------------------------------------------------------------------------
# parameter construction (shapes/names of the reference state_dict) + search step
# --------------------------------------------------------------------------
def arch_shapes(cfg):
    k_a = sum(cfg.num_input_nodes + i for i in range(cfg.steps))
    k_b = sum(2 + i for i in range(cfg.node_steps))
    shapes = [(k_a, 2)]
    for _ in range(cfg.steps):
        shapes += [(k_b, 2), (cfg.node_steps, len(cfg.step_ops))]
    return shapes


def param_shapes(cfg, num_classes=None, prefix='cell', genotype=None):
    """name -> (shape, kind) in the reference's state_dict order.  kind in
    {'w' weight/trainable, 'rm','rv','nbt' BatchNorm buffers}."""
    C, L = cfg.C, cfg.L
    out = {}

    def bn(p, n):
        out[p + '.weight'] = ((n,), 'w')
        out[p + '.bias'] = ((n,), 'w')
        out[p + '.running_mean'] = ((n,), 'rm')
        out[p + '.running_var'] = ((n,), 'rv')
        out[p + '.num_batches_tracked'] = ((), 'nbt')

    def op(p, name):
        if name == 'ScaleDotAttn':
            out[p + '.ln.weight'] = ((C, L), 'w')
            out[p + '.ln.bias'] = ((C, L), 'w')
        elif name == 'LinearGLU':
            out[p + '.conv.weight'] = ((2 * C, 2 * C, 1), 'w')
            out[p + '.conv.bias'] = ((2 * C,), 'w')
            bn(p + '.bn', 2 * C)
        elif name in ('ConcatFC', 'CatConvMish'):
            out[p + '.conv.weight'] = ((C, 2 * C, 1), 'w')
            out[p + '.conv.bias'] = ((C,), 'w')
            bn(p + '.bn', C)

    mult = cfg.multiplier if genotype is None else len(genotype.concat)
    steps = cfg.steps if genotype is None else len(genotype.edges) // 2
    for i in range(steps):
        nc = f'{prefix}._step_nodes.{i}.node_cell'
        for j in range(cfg.node_steps):
            if genotype is None:
                for k, name in enumerate(cfg.step_ops):
                    op(f'{nc}.node_ops.{j}._ops.{k}', name)
            else:
                op(f'{nc}.node_ops.{j}', genotype.steps[i].inner_steps[j])
        if cfg.node_multiplier != 1:
            out[nc + '.out_conv.weight'] = ((C, C * cfg.node_multiplier, 1), 'w')
            out[nc + '.out_conv.bias'] = ((C,), 'w')
            bn(nc + '.bn', C)
        out[nc + '.ln.weight'] = ((C, L), 'w')
        out[nc + '.ln.bias'] = ((C, L), 'w')
    out[prefix + '.ln.weight'] = ((C * mult, L), 'w')   # registered after the step nodes
    out[prefix + '.ln.bias'] = ((C * mult, L), 'w')
    if num_classes is not None:
        out = {('fusion_net.' + k): v for k, v in out.items()}
        out['central_classifier.weight'] = ((num_classes, C * L * mult), 'w')
        out['central_classifier.bias'] = ((num_classes,), 'w')
    return out


def init_params(cfg, num_classes=None, seed=0, dtype=torch.float32, prefix='cell', genotype=None):
    """Random but well-conditioned parameters with the reference's names/shapes
    (NOT the reference's init RNG stream -- parity tests copy tensors across)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, (shape, kind) in param_shapes(cfg, num_classes, prefix, genotype).items():
        if kind == 'nbt':
            P[name] = torch.zeros((), dtype=torch.int64)
        elif kind == 'rm':
            P[name] = torch.zeros(shape, dtype=dtype)
        elif kind == 'rv':
            P[name] = torch.ones(shape, dtype=dtype)
        elif name.endswith('conv.weight') or name.endswith('classifier.weight'):
            fan_in = shape[1]
            P[name] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul(1 / math.sqrt(fan_in)).to(dtype)
        elif name.endswith('conv.bias') or name.endswith('classifier.bias'):
            P[name] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul(0.05).to(dtype)
        elif name.endswith('.weight'):   # BN / LN gains: around 1
            P[name] = (1 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)).to(dtype)
        else:                            # BN / LN biases: around 0
            P[name] = (0.1 * torch.randn(shape, generator=g, dtype=torch.float64)).to(dtype)
    return P


def init_arch(cfg, seed=0, scale=1e-3, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed + 1000)
    return [(scale * torch.randn(s, generator=g, dtype=torch.float64)).to(dtype) for s in arch_shapes(cfg)]


def trainable_names(P):
    return [k for k in P if not (k.endswith('running_mean') or k.endswith('running_var')
                                 or k.endswith('num_batches_tracked'))]


def loss_and_grads(feats, labels, arch, P, masks, cfg, loss='ce', genotype=None, training=True):
    """One fwd+bwd of the head: returns (loss, logits, weight grads dict, arch grads list)."""
    names = trainable_names(P)
    leaves = {k: P[k].detach().clone().requires_grad_(True) for k in names}
    Pl = dict(P)
    Pl.update(leaves)
    al = [a.detach().clone().requires_grad_(True) for a in arch] if arch is not None else None
    logits = head_logits(feats, al, Pl, masks, training, cfg, genotype)
    if loss == 'ce':
        lv = F.cross_entropy(logits, labels)
    else:
        lv = F.binary_cross_entropy_with_logits(logits, labels)
    wrt = [leaves[k] for k in names] + (al if al is not None else [])
    gs = torch.autograd.grad(lv, wrt, allow_unused=True)
    gw = {k: g for k, g in zip(names, gs[:len(names)])}
    ga = list(gs[len(names):])
    # BN buffers were updated in place on the shared P tensors
    return lv.detach(), logits.detach(), gw, ga


def unrolled_arch_grad(train, valid, arch, P, cfg, eta, weight_decay, masks=None, loss='ce'):
    """Second-order DARTS architecture gradient (Liu et al. 2019, eq. 7-8, finite-difference Hessian-vector product) --
    the update the reference's unused ``--unrolled`` flag (main_darts_found_ntu.py:48) names; restated here so that the
    product's Architect.step_unrolled has an independent CPU check.  train / valid = (feats, labels)."""
    names = trainable_names(P)
    Pw = {k: v.clone() for k, v in P.items()}
    _, _, gw, _ = loss_and_grads(train[0], train[1], arch, Pw, masks, cfg, loss)
    P1 = {k: v.clone() for k, v in P.items()}
    for k in names:
        P1[k] = P[k] - eta * (gw[k] + weight_decay * P[k])
    _, _, dw, da = loss_and_grads(valid[0], valid[1], arch, P1, masks, cfg, loss)
    norm = math.sqrt(sum(float((dw[k].double() ** 2).sum()) for k in names))
    eps = 0.01 / max(norm, 1e-30)
    outs = []
    for sign in (1.0, -1.0):
        Pp = {k: v.clone() for k, v in P.items()}
        for k in names:
            Pp[k] = P[k] + sign * eps * dw[k]
        outs.append(loss_and_grads(train[0], train[1], arch, Pp, masks, cfg, loss)[3])
    return [d - eta * (p - n) / (2.0 * eps) for d, p, n in zip(da, outs[0], outs[1])]


class SearchState:
    """Everything one search run carries: weights, arch tensors, both Adam states,
    LR schedule.  ``search_step`` restates train_searchable/ntu.py:70-93 +
    architect.py:21-29 for one (dev batch, train batch) pair."""

    def __init__(self, cfg, P, arch, eta_max=1e-3, eta_min=1e-6, Ti=1, Tm=2, nbpe=100,
                 weight_decay=3e-4, arch_lr=3e-4, arch_wd=1e-3, loss='ce'):
        self.cfg, self.P, self.arch = cfg, P, arch
        self.sched = CosineRestartLR(eta_max, eta_min, Ti, Tm, nbpe)
        self.wd, self.arch_lr, self.arch_wd, self.loss = weight_decay, arch_lr, arch_wd, loss
        self.w_state, self.a_state = {}, {}
        self.names = trainable_names(P)

    def arch_step(self, feats, labels, masks=None):
        lv, logits, gw, ga = loss_and_grads(feats, labels, self.arch, self.P, masks, self.cfg, self.loss)
        adam_step(self.arch, ga, self.a_state, self.arch_lr, (0.5, 0.999), self.arch_wd)
        return lv

    def weight_step(self, feats, labels, masks=None):
        lv, logits, gw, ga = loss_and_grads(feats, labels, self.arch, self.P, masks, self.cfg, self.loss)
        lr = self.sched.step()
        adam_step([self.P[k] for k in self.names], [gw[k] for k in self.names], self.w_state,
                  lr, (0.9, 0.999), self.wd)
        return lv

    def search_step(self, dev, train, masks_dev=None, masks_train=None):
        la = self.arch_step(dev[0], dev[1], masks_dev)
        lw = self.weight_step(train[0], train[1], masks_train)
        return la, lw

    def genotype(self):
        return network_genotype(self.arch, self.cfg)


def synthetic_batch(cfg, B, num_classes, seed=2, loss='ce', dtype=torch.float32):
    """SURVEY 8d synthetic inputs: unit-normal (B,C,L) features, one generator per node."""
    feats = [torch.randn(B, cfg.C, cfg.L, generator=torch.Generator().manual_seed(seed + i)).to(dtype)
             for i in range(cfg.num_input_nodes)]
    g = torch.Generator().manual_seed(seed + 100)
    if loss == 'ce':
        labels = torch.randint(0, num_classes, (B,), generator=g)
    else:
        labels = (torch.rand(B, num_classes, generator=g) < 0.2).to(dtype)
    return feats, labels


def dump_genotype(genotype) -> bytes:
    """The pickle the reference writes (darts/utils.py:96-99) resolves these
    namedtuples at ``models.search.darts.genotypes``; callers patch __module__."""
    return pickle.dumps(genotype)
