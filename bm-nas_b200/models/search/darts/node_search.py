"""Lower-level searchable cell -- drop-in for models/search/darts/node_search.py
(NodeCell :12-70, FusionNode :72-163).

``FusionNode.forward`` hands the raw beta/gamma logits to the kernels (the per-edge
2-way softmax and the per-step softmax over primitives are taken in-kernel, and the
backward returns d/d(logits) directly); ``NodeCell.forward`` keeps the reference
signature and accepts already soft-maxed weights.  beta/gamma are plain tensors with
``requires_grad`` (not nn.Parameters, not in ``state_dict`` -- SURVEY fact 4) but, unlike
the reference, they follow the module in ``.to(device)``.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from bmnas import runtime as _rt

from . import genotypes as _gt
from .genotypes import STEP_EDGE_PRIMITIVES, StepGenotype
from .node_operations import NodeMixedOp, _dropkw
from .operations import FusionMixedOp


class NodeCell(nn.Module):
    def __init__(self, node_steps, node_multiplier, args):
        super().__init__()
        self.args = args
        self.node_steps = node_steps
        self.node_multiplier = node_multiplier
        self.C, self.L = args.C, args.L
        self.num_input_nodes = 2
        self.edge_ops = nn.ModuleList()
        self.node_ops = nn.ModuleList()
        for i in range(node_steps):
            for _ in range(self.num_input_nodes + i):
                self.edge_ops.append(FusionMixedOp(self.C, self.L, args))
        for i in range(node_steps):
            self.node_ops.append(NodeMixedOp(self.C, self.L, args))
        if node_multiplier != 1:
            self.out_conv = nn.Conv1d(self.C * node_multiplier, self.C, 1, 1)
            self.bn = nn.BatchNorm1d(self.C)
            self.out_dropout = nn.Dropout(args.drpt)
        self.ln = nn.LayerNorm([self.C, self.L])
        self.dropout = nn.Dropout(args.drpt)     # constructed but unused, as in the reference (C-5)

    def _op_names(self):
        return list(self.node_ops[0]._names)

    def _run(self, owner, x, y, edge_w, node_w, logits, arch_leaves):
        """shared by NodeCell.forward (weights given) and FusionNode.forward (logits)."""
        B, C, L = x.shape
        alias = x is y
        ops, ns, nm = self._op_names(), self.node_steps, self.node_multiplier
        P = _rt.named_tensors(self, prefix='node_cell.')
        n_x = 1 if alias else 2

        def build(prog, slots, need, G):
            G.attach(P)
            out = prog.buf(B, C, L)
            xs = slots[0]
            ys = slots[0] if alias else slots[1]
            if logits:
                ew, nw = edge_w, node_w
                gew, gnw = G.of(edge_w), G.of(node_w)
            else:
                ew, nw = slots[n_x], slots[n_x + 1]
                gew = prog.buf(*edge_w.shape) if need[n_x] else None
                gnw = prog.buf(*node_w.shape) if need[n_x + 1] else None
                if gew is not None:
                    prog.out_grad(ew, gew)
                if gnw is not None:
                    prog.out_grad(nw, gnw)
            prog.node_cell_search(xs, ys, need[0], need[0] if alias else need[1], ew, nw, logits, gew, gnw, P, G,
                                  'node_cell', ops, ns, nm, out)
            return out
        ins = [x] if alias else [x, y]
        if not logits:
            ins += [edge_w, node_w]
        leaves = list(self.parameters()) + list(arch_leaves)
        return _rt.run(owner, 'node_search', ins, build, leaves, C, L, self.args.drpt,
                       key_extra=(alias, logits, tuple(ops)), **_dropkw(self, 'node_cell'))

    def forward(self, x, y, edge_weights, node_weights):
        return self._run(self, x, y, edge_weights, node_weights, False, [])


class FusionNode(nn.Module):
    def __init__(self, node_steps, node_multiplier, args):
        super().__init__()
        self.node_steps = node_steps
        self.node_multiplier = node_multiplier
        self.node_cell = NodeCell(node_steps, node_multiplier, args)
        self.num_input_nodes = 2
        self.num_keep_edges = 2
        self._initialize_betas()
        self._initialize_gammas()
        self._arch_parameters = [self.betas, self.gammas]

    def _initialize_betas(self):
        k = sum(self.num_input_nodes + i for i in range(self.node_steps))
        self.betas = (1e-3 * torch.randn(k, len(STEP_EDGE_PRIMITIVES))).requires_grad_(True)

    def _initialize_gammas(self):
        self.gammas = (1e-3 * torch.randn(self.node_steps, len(self.node_cell._op_names()))).requires_grad_(True)

    def _apply(self, fn, *a, **kw):
        super()._apply(fn, *a, **kw)
        for t in self._arch_parameters:       # same tensor objects stay registered with the arch optimiser
            t.data = fn(t.data)
        return self

    def forward(self, x, y):
        return self.node_cell._run(self, x, y, self.betas, self.gammas, True, self._arch_parameters)

    def arch_parameters(self):
        return self._arch_parameters

    def node_genotype(self):
        """argmax derivation, identical tie-breaking to node_search.py:110-163: top-2 predecessors by
        their non-'none' weight (stable sort), best non-'none' op per edge and best primitive per inner
        step with strict '>' (ties -> lowest index)."""
        ew = F.softmax(self.betas.detach().float().cpu(), dim=-1)
        nw = F.softmax(self.gammas.detach().float().cpu(), dim=-1)
        none = STEP_EDGE_PRIMITIVES.index('none')
        names = self.node_cell._op_names()
        edge_gene, node_gene = [], []
        start = 0
        for i in range(self.node_steps):
            n = self.num_input_nodes + i
            W = ew[start:start + n]
            strength = [max(W[r][k] for k in range(W.shape[1]) if k != none) for r in range(n)]
            keep = sorted(range(n), key=lambda r: -strength[r])[:self.num_keep_edges]
            for j in keep:
                best = None
                for k in range(W.shape[1]):
                    if k != none and (best is None or W[j][k] > W[j][best]):
                        best = k
                edge_gene.append((STEP_EDGE_PRIMITIVES[best], j))
            start += n
        for i in range(self.node_steps):
            best = None
            for k in range(nw.shape[1]):
                if best is None or nw[i][k] > nw[i][best]:
                    best = k
            node_gene.append(names[best])
        concat = list(range(self.num_input_nodes + self.node_steps - self.node_multiplier,
                            self.node_steps + self.num_input_nodes))
        return StepGenotype(inner_edges=edge_gene, inner_steps=node_gene, inner_concat=concat)
