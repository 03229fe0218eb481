"""Fixed-genotype ("found") lower cell -- drop-in for models/search/darts/node.py
(Found_NodeCell :8-76, Found_FusionNode :78-91).  One concrete primitive per inner
step, x != y in general; same fused kernels as the search path with weight 1 per op.
The reference's dead ablation nodes (Found_DARTS_/MFAS_/AOA_/TwoHeadAttn_FusionNode,
node.py:94-183, two of which raise KeyError as shipped -- SURVEY C-7) are not provided.
"""
import torch.nn as nn

from bmnas import runtime as _rt

from .node_operations import STEP_STEP_OPS, _dropkw
from .operations import OPS


class Found_NodeCell(nn.Module):
    def __init__(self, node_steps, node_multiplier, args, step_genotype):
        super().__init__()
        self.args = args
        self.node_steps = node_steps
        self.node_multiplier = node_multiplier
        self.C, self.L = args.C, args.L
        self.num_input_nodes = 2
        self.edge_ops = nn.ModuleList()
        self.node_ops = nn.ModuleList()
        self._gene = step_genotype
        op_names, indices = zip(*step_genotype.inner_edges)
        self.compile(op_names, indices, step_genotype.inner_steps)
        if node_multiplier != 1:
            self.out_conv = nn.Conv1d(self.C * node_multiplier, self.C, 1, 1)
            self.bn = nn.BatchNorm1d(self.C)
            self.out_dropout = nn.Dropout(args.drpt)
        self.ln = nn.LayerNorm([self.C, self.L])
        self.dropout = nn.Dropout(args.drpt)

    def compile(self, edge_op_names, edge_indices, inner_steps):
        for name in edge_op_names:
            self.edge_ops.append(OPS[name](self.C, self.L, self.args))
        self.edge_indices = edge_indices
        for name in inner_steps:
            self.node_ops.append(STEP_STEP_OPS[name](self.C, self.L, self.args))

    def _run(self, owner, x, y):
        B, C, L = x.shape
        alias = x is y
        gene, ns, nm = self._gene, self.node_steps, self.node_multiplier
        P = _rt.named_tensors(self, prefix='node_cell.')

        def build(prog, slots, need, G):
            G.attach(P)
            out = prog.buf(B, C, L)
            xs = slots[0]
            ys = slots[0] if alias else slots[1]
            prog.node_cell_found(xs, ys, need[0], need[0] if alias else need[1], gene, P, G, 'node_cell', ns, nm, out)
            return out
        return _rt.run(owner, 'node_found', [x] if alias else [x, y], build, list(self.parameters()), C, L,
                       self.args.drpt, key_extra=(alias,), **_dropkw(self, 'node_cell'))

    def forward(self, x, y):
        return self._run(self, x, y)


class Found_FusionNode(nn.Module):
    def __init__(self, node_steps, node_multiplier, args, step_genotype):
        super().__init__()
        self.node_steps = node_steps
        self.node_multiplier = node_multiplier
        self.node_cell = Found_NodeCell(node_steps, node_multiplier, args, step_genotype)
        self.num_input_nodes = 2
        self.num_keep_edges = 2

    def forward(self, x, y):
        return self.node_cell._run(self, x, y)
