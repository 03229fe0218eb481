"""Reshape layers right upstream of the fusion cells -- drop-in for ReshapeInputLayer (:51-76) and
ReshapeInputLayer_MMIMDB (:87-115) of models/auxiliary/aux_models.py.

A raw backbone feature (B, C_in), (B, C_in, d2) or (B, C_in, d2, ...) becomes the (B, C, L) tensor the
FusionNetwork consumes:  adaptive max pool -> [F.interpolate(size=L): the identity] -> Conv1d(C_in, C, 1) ->
BatchNorm1d(C) -> ReLU -> Dropout(args.drpt).  The nn.Conv1d / nn.BatchNorm1d / nn.Dropout children only hold
parameters and buffers (same names, shapes and initialisation as the reference, so ``state_dict`` round-trips);
the arithmetic is three launches: bmnas_pool_fwd (csrc/pool.cu), bmnas_conv_fwd (GEMM over the large C_in with
the BatchNorm statistics in its epilogue) and bmnas_node_fwd (BN apply + ReLU + dropout), and their backward.
"""
import math

import torch.nn as nn

from bmnas import runtime as _rt
from models.search.darts.node_operations import _dropkw


class _Reshape(nn.Module):
    def __init__(self, C_in, C, L, args):
        super().__init__()
        self.C = C
        self.L = L
        self.conv = nn.Conv1d(C_in, self.C, 1, 1)
        self.bn = nn.BatchNorm1d(self.C)
        self.dropout = nn.Dropout(args.drpt)

    def _bins(self):
        raise NotImplementedError

    def forward(self, x):
        B, Cin = x.shape[0], x.shape[1]
        if Cin != self.conv.in_channels:
            raise RuntimeError(f'expected {self.conv.in_channels} input channels, got {Cin}')
        # x.unsqueeze(-1) [.unsqueeze(-1)] .view(B, C_in, size(2), -1)   (aux_models.py:62-63, 103-106)
        H = x.shape[2] if x.dim() > 2 else 1
        W = x.numel() // (B * Cin * H)
        x = x.reshape(B, Cin, H, W)          # the view autograd maps the (B, C_in, H, W) gradient back through
        OH, OW = self._bins()
        C, L = self.C, self.L
        P = _rt.named_tensors(self, prefix='op.')

        def build(prog, slots, need, G):
            G.attach(P)
            out = prog.buf(B, C, L)
            pooled = prog.pool(slots[0], Cin, H, W, OH, OW, need[0])
            prog.node_op(None, None, ['ConcatFC'], P, G, (lambda k: 'op'), None, 0, False, out,
                         conv_srcs=[pooled], conv_src_C=[Cin], conv_need=[need[0]])
            return out
        return _rt.run(self, 'reshape', [x], build, list(self.parameters()), C, L, self.dropout.p,
                       key_extra=(tuple(x.shape), OH, OW), **_dropkw(self, 'op'))


class ReshapeInputLayer(_Reshape):
    """NTU / EgoGesture: pool dim 2 onto L bins and everything behind it onto one (aux_models.py:58,66)."""

    def __init__(self, C_in, C, L, args):
        super().__init__(C_in, C, L, args)
        self.pool = nn.AdaptiveMaxPool2d((self.L, 1))      # attribute kept for parity with the reference module tree

    def _bins(self):
        return self.L, 1


class ReshapeInputLayer_MMIMDB(_Reshape):
    """MM-IMDB: pool the feature map onto a sqrt(L) x sqrt(L) grid (aux_models.py:95-97,108)."""

    def __init__(self, C_in, C, L, args):
        super().__init__(C_in, C, L, args)
        pool_size = int(math.sqrt(self.L * 1.0))
        assert pool_size * pool_size == self.L
        self.pool = nn.AdaptiveMaxPool2d((pool_size, pool_size))
        self._ps = pool_size

    def _bins(self):
        return self._ps, self._ps
