"""Fixed-genotype ("found") upper cell and network -- drop-in for models/search/darts/model.py
(Found_FusionCell :16-89, Found_Random_FusionCell :92-160 -- textually identical in the
reference, C-6 -- and Found_FusionNetwork :162-190).  The whole found network runs as one
static launch plan (forward and backward).
"""
import torch.nn as nn

from bmnas import runtime as _rt

from .node import Found_FusionNode
from .node_operations import _dropkw
from .operations import OPS


class Found_FusionCell(nn.Module):
    def __init__(self, steps, args, genotype):
        super().__init__()
        self.C, self.L = args.C, args.L
        self.args = args
        self._genotype = genotype
        op_names, indices = zip(*genotype.edges)
        self._compile(self.C, self.L, op_names, indices, genotype.concat, genotype.steps, args)
        self._steps = steps
        self.ln = nn.LayerNorm([self.C * self._multiplier, self.L])

    def _compile(self, C, L, op_names, indices, concat, gene_step_nodes, args):
        assert len(op_names) == len(indices)
        self._steps = len(op_names) // 2
        self._concat = concat
        self._multiplier = len(concat)
        self._ops = nn.ModuleList(OPS[name](C, L, args) for name in op_names)
        self._indices = indices
        self._step_nodes = nn.ModuleList(
            Found_FusionNode(args.node_steps, args.node_multiplier, args, g) for g in gene_step_nodes)

    def _run(self, owner, prefix, feats):
        B, C, L = feats[0].shape
        args, gt = self.args, self._genotype
        if self._steps != len(gt.edges) // 2:
            # the reference overwrites _steps with the constructor argument after _compile (model.py:105)
            # and then indexes past the genotype; fail loudly instead.
            raise ValueError(f'steps={self._steps} does not match the genotype ({len(gt.edges) // 2} steps)')
        P = _rt.named_tensors(self, prefix=prefix + '.')

        def build(prog, slots, need, G):
            G.attach(P)
            return prog.cell_found(slots, list(need), gt, P, G, prefix, args.node_steps, args.node_multiplier)
        out = _rt.run(owner, 'cell_found', list(feats), build, list(self.parameters()), C, L, args.drpt,
                      key_extra=(len(feats),), **_dropkw(self, prefix))
        return out.view(B, -1)

    def forward(self, input_features):
        return self._run(self, 'cell', list(input_features))


class Found_Random_FusionCell(Found_FusionCell):
    pass


class Found_FusionNetwork(nn.Module):
    def __init__(self, steps, multiplier, num_input_nodes, num_keep_edges, args, criterion, genotype):
        super().__init__()
        self._steps = steps
        self._multiplier = multiplier
        self._criterion = criterion
        self._genotype = genotype
        self._num_input_nodes = num_input_nodes
        self._num_keep_edges = num_keep_edges
        self.cell = Found_Random_FusionCell(steps, args, genotype)

    def forward(self, input_features):
        assert self._num_input_nodes == len(input_features)
        return self.cell._run(self, 'cell', list(input_features))

    def _loss(self, input_features, labels):
        return self._criterion(self(input_features), labels)

    def get_genotype(self):
        return self._genotype
