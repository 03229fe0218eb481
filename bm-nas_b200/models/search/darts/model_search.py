"""Upper-level searchable cell and network -- drop-in for models/search/darts/model_search.py
(FusionCell :13-68, FusionNetwork :70-181).

``FusionNetwork.forward`` executes ONE static launch plan for the whole hypernet (edge
mixes, every step node, cell tail) forward and backward; alpha/beta/gamma logits go to
the kernels raw.  ``FusionCell.forward(input_features, weights)`` keeps the reference
signature (soft-maxed alpha weights passed in).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from bmnas import runtime as _rt

from .genotypes import PRIMITIVES, Genotype
from .node_operations import _dropkw
from .node_search import FusionNode
from .operations import FusionMixedOp


class FusionCell(nn.Module):
    def __init__(self, steps, multiplier, args):
        super().__init__()
        self._steps = steps
        self._multiplier = multiplier
        self.args = args
        self._ops = nn.ModuleList()
        self._step_nodes = nn.ModuleList()
        self.num_input_nodes = args.num_input_nodes
        self.C, self.L = args.C, args.L
        self.ln = nn.LayerNorm([self.C * multiplier, self.L])
        for i in range(steps):
            for _ in range(self.num_input_nodes + i):
                self._ops.append(FusionMixedOp(self.C, self.L, args))
        self._initialize_step_nodes(args)

    def _initialize_step_nodes(self, args):
        for _ in range(self._steps):
            self._step_nodes.append(FusionNode(args.node_steps, args.node_multiplier, args))

    def arch_parameters(self):
        self._arch_parameters = []
        for node in self._step_nodes:
            self._arch_parameters += node.arch_parameters()
        return self._arch_parameters

    def _run(self, owner, prefix, feats, alphas, alpha_logits, extra_leaves):
        B, C, L = feats[0].shape
        n_in = len(feats)
        args = self.args
        ops = self._step_nodes[0].node_cell._op_names()
        P = _rt.named_tensors(self, prefix=prefix + '.')
        node_arch = [(n.betas, n.gammas) for n in self._step_nodes]

        def build(prog, slots, need, G):
            G.attach(P)
            if alpha_logits:
                al, gal = alphas, G.of(alphas)
            else:
                al = slots[n_in]
                gal = prog.buf(*alphas.shape) if need[n_in] else None
                if gal is not None:
                    prog.out_grad(al, gal)
            g_node = [(G.of(b), G.of(g)) for b, g in node_arch]
            out = prog.cell_search(slots[:n_in], list(need[:n_in]), al, gal, node_arch, g_node, alpha_logits, P, G,
                                   prefix, self._steps, self._multiplier, ops, args.node_steps, args.node_multiplier)
            return out
        ins = list(feats) + ([] if alpha_logits else [alphas])
        leaves = list(self.parameters()) + self.arch_parameters() + list(extra_leaves)
        out = _rt.run(owner, 'cell_search', ins, build, leaves, C, L, args.drpt,
                      key_extra=(n_in, alpha_logits, tuple(ops)), **_dropkw(self, prefix))
        return out.view(B, -1)

    def forward(self, input_features, weights):
        return self._run(self, 'cell', list(input_features), weights, False, [])


class FusionNetwork(nn.Module):
    def __init__(self, steps, multiplier, num_input_nodes, num_keep_edges, args, criterion=None, logger=None):
        super().__init__()
        self.logger = logger
        self._steps = steps
        self._multiplier = multiplier
        self._criterion = criterion
        self._num_input_nodes = num_input_nodes
        self._num_keep_edges = num_keep_edges
        self.cell = FusionCell(steps, multiplier, args)
        self.cell_arch_parameters = self.cell.arch_parameters()
        self._initialize_alphas()
        self._arch_parameters = [self.alphas_edges] + self.cell_arch_parameters

    def _initialize_alphas(self):
        k = sum(self._num_input_nodes + i for i in range(self._steps))
        self.alphas_edges = (1e-3 * torch.randn(k, len(PRIMITIVES))).requires_grad_(True)

    def _apply(self, fn, *a, **kw):
        super()._apply(fn, *a, **kw)
        self.alphas_edges.data = fn(self.alphas_edges.data)
        return self

    def forward(self, input_features):
        assert self._num_input_nodes == len(input_features)
        return self.cell._run(self, 'cell', list(input_features), self.alphas_edges, True, [self.alphas_edges])

    def _loss(self, input_features, labels):
        return self._criterion(self(input_features), labels)

    def arch_parameters(self):
        return self._arch_parameters

    def genotype(self):
        """'sample strategy v3' of model_search.py:111-181: per step, the best-scoring pair of ORIGINAL
        input nodes with at least one not used yet (score = product of their non-'none' weights; ties ->
        first pair in lexicographic order), then the best non-'none' op per chosen edge."""
        W_all = F.softmax(self.alphas_edges.detach().float(), dim=-1).cpu().numpy()
        none = PRIMITIVES.index('none')
        n_in = self._num_input_nodes
        gene, used = [], []
        start = 0
        for i in range(self._steps):
            n = n_in + i
            W = W_all[start:start + n].copy()
            strength = [max(W[j][t] for t in range(W.shape[1]) if t != none) for j in range(n_in)]
            pairs = [[j, k, strength[j] * strength[k]] for j in range(n_in) for k in range(j + 1, n_in)
                     if (j not in used) or (k not in used)]
            j, k, _ = sorted(pairs, key=lambda pr: -pr[2])[0]
            used = list(set(used + [j, k]))
            for e in (j, k):
                best = None
                for t in range(W.shape[1]):
                    if t != none and (best is None or W[e][t] > W[e][best]):
                        best = t
                gene.append((PRIMITIVES[best], e))
            start += n
        steps = [node.node_genotype() for node in self.cell._step_nodes]
        concat = list(range(n_in + self._steps - self._multiplier, self._steps + n_in))
        return Genotype(edges=gene, concat=concat, steps=steps)
