"""Edge primitives and the edge MixedOp -- drop-in for models/search/darts/operations.py.

Registry and classes keep the reference names (OPS :7-12, Zero :14-20, Identity :88-93,
FusionMixedOp :95-106).  With PRIMITIVES = ['none', 'skip'] a FusionMixedOp is exactly
``w[skip] * x``; its forward/backward run in the ``bmnas_mix_*`` CUDA
kernels.  The stand-alone ``fc_relu`` / ``fc_mish`` candidates of the reference registry
are not in PRIMITIVES and are not part of the search-step path (out of scope).
"""
import torch
import torch.nn as nn

from bmnas import runtime as _rt

from .genotypes import PRIMITIVES


class Zero(nn.Module):
    """'none': x * 0 (finite inputs)."""

    def forward(self, x):
        return x.mul(0.)


class Identity(nn.Module):
    """'skip'."""

    def forward(self, x):
        return x


OPS = {
    'none': lambda C, L, args: Zero(),
    'skip': lambda C, L, args: Identity(),
}


class FusionMixedOp(nn.Module):
    """forward(x, weights[len(PRIMITIVES)]) = sum_k weights[k] * op_k(x)."""

    def __init__(self, C, L, args):
        super().__init__()
        self._ops = nn.ModuleList(OPS[p](C, L, args) for p in PRIMITIVES)
        self._C, self._L = C, L

    def forward(self, x, weights):
        B, C, L = x.shape

        def build(prog, slots, need, G):
            out = prog.buf(B, C, L)
            gw = prog.buf(2) if need[1] else None
            if gw is not None:
                prog.out_grad(slots[1], gw)
            prog.mix([slots[0]], slots[1], 0, False, out, gw=gw, need=[need[0]])
            return out
        return _rt.run(self, 'edge', [x, weights], build, [], C, L, 0.0)
