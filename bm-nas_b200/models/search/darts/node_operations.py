"""Step-node fusion primitives and the node MixedOp -- drop-in for
models/search/darts/node_operations.py (STEP_STEP_OPS :9-14, Sum :16-20, LinearGLU :22-39,
ConcatFC :41-56, CatConvMish :66-82, ScaledDotAttn :84-108, NodeMixedOp :110-120).

The nn.Conv1d / nn.BatchNorm1d / nn.LayerNorm / nn.Dropout children only hold the
parameters and buffers (identical names, shapes and default initialisation as the
reference, so ``state_dict`` round-trips); the arithmetic runs in the fused CUDA kernels
``bmnas_conv_*`` (GEMM + BN statistics) and ``bmnas_node_*`` (all primitives + the
weighted sum).  Parity-mode dropout: set ``module.dropout.injected_mask`` (uint8 keep
mask); otherwise masks come from the in-kernel Philox stream.
"""
import torch
import torch.nn as nn

from bmnas import runtime as _rt

from .genotypes import STEP_STEP_PRIMITIVES

STEP_STEP_OPS = {
    'Sum': lambda C, L, args: Sum(),
    'ScaleDotAttn': lambda C, L, args: ScaledDotAttn(C, L),
    'LinearGLU': lambda C, L, args: LinearGLU(C, args),
    'ConcatFC': lambda C, L, args: ConcatFC(C, args),
}


def collect_dropout(root, prefix=''):
    """({site: injected uint8 keep mask} or None, {site: p}) over every nn.Dropout below root.
    The probabilities are read from the modules, so ``module.dropout.p = ...`` behaves as in torch."""
    masks, ps = {}, {}
    for name, m in root.named_modules(prefix=prefix):
        if isinstance(m, nn.Dropout):
            ps[name] = float(m.p)
            im = getattr(m, 'injected_mask', None)
            if im is not None:
                masks[name] = im
    return (masks or None), ps


def _dropkw(root, prefix):
    masks, ps = collect_dropout(root, prefix)
    return dict(masks=masks, drop_p=ps)


class _Primitive(nn.Module):
    """a single step-node primitive evaluated by the fused node kernel with weight 1"""
    _bm_name = None

    def forward(self, x, y):
        B, C, L = x.shape
        alias = x is y
        name = self._bm_name
        P = _rt.named_tensors(self, prefix='op.')
        drpt = getattr(getattr(self, 'dropout', None), 'p', 0.0) if name != 'ScaleDotAttn' else 0.0

        def build(prog, slots, need, G):
            G.attach(P)
            out = prog.buf(B, C, L)
            xs = slots[0]
            ys = slots[0] if alias else slots[1]
            prog.node_op(xs, ys, [name], P, G, (lambda k: 'op'), None, 0, False, out,
                         need_x=need[0], need_y=need[0] if alias else need[1])
            return out
        leaves = list(self.parameters())
        return _rt.run(self, 'prim', [x] if alias else [x, y], build, leaves, C, L, drpt,
                       key_extra=(alias,), **_dropkw(self, 'op'))


class Sum(_Primitive):
    _bm_name = 'Sum'

    def __init__(self):
        super().__init__()


class LinearGLU(_Primitive):
    _bm_name = 'LinearGLU'

    def __init__(self, C, args):
        super().__init__()
        self.conv = nn.Conv1d(2 * C, 2 * C, 1, 1)
        self.bn = nn.BatchNorm1d(2 * C)
        self.dropout = nn.Dropout(args.drpt)


class ConcatFC(_Primitive):
    _bm_name = 'ConcatFC'

    def __init__(self, C, args):
        super().__init__()
        self.conv = nn.Conv1d(2 * C, C, 1, 1)
        self.bn = nn.BatchNorm1d(C)
        self.dropout = nn.Dropout(args.drpt)


class Mish(nn.Module):
    def forward(self, x):
        return x * torch.tanh(nn.functional.softplus(x))


class CatConvMish(_Primitive):
    """not registered in STEP_STEP_OPS by default (as in the reference); append it to
    STEP_STEP_OPS / STEP_STEP_PRIMITIVES at run time to search over it."""
    _bm_name = 'CatConvMish'

    def __init__(self, C, args):
        super().__init__()
        self.conv = nn.Conv1d(2 * C, C, 1, 1)
        self.bn = nn.BatchNorm1d(C)
        self.dropout = nn.Dropout(args.drpt)
        self.mish = Mish()


class ScaledDotAttn(_Primitive):
    """single-head L x L attention of x over y with head dim C, dropout(0.1), LayerNorm([C, L])"""
    _bm_name = 'ScaleDotAttn'

    def __init__(self, C, L):
        super().__init__()
        self.dropout = nn.Dropout(0.1)
        self.ln = nn.LayerNorm([C, L])


class NodeMixedOp(nn.Module):
    """forward(x, y, weights[n_ops]) = sum_k weights[k] * op_k(x, y) in ONE fused kernel."""

    def __init__(self, C, L, args):
        super().__init__()
        self._ops = nn.ModuleList(STEP_STEP_OPS[p](C, L, args) for p in STEP_STEP_PRIMITIVES)
        self._names = list(STEP_STEP_PRIMITIVES)
        self._drpt = args.drpt

    def forward(self, x, y, weights):
        B, C, L = x.shape
        alias = x is y
        names = self._names
        P = _rt.named_tensors(self, prefix='mix.')

        def build(prog, slots, need, G):
            G.attach(P)
            out = prog.buf(B, C, L)
            xs = slots[0]
            ys = slots[0] if alias else slots[1]
            ws = slots[-1]
            gw = prog.buf(len(names)) if need[-1] else None
            if gw is not None:
                prog.out_grad(ws, gw)
            prog.node_op(xs, ys, names, P, G, (lambda k: f'mix._ops.{k}'), ws, 0, False, out, g_gamma=gw,
                         need_x=need[0], need_y=need[0] if alias else need[1])
            return out
        ins = [x, weights] if alias else [x, y, weights]
        return _rt.run(self, 'mixed', ins, build, list(self.parameters()), C, L, self._drpt,
                       key_extra=(alias, tuple(names)), **_dropkw(self, 'mix'))
