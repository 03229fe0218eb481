"""Genotype format of BM-NAS (drop-in for models/search/darts/genotypes.py:3-21).

The namedtuple names, field order and this module path are part of the on-disk
contract: ``best_genotype.pkl`` files written by the reference unpickle against
``models.search.darts.genotypes``, and files written here load in
the reference.

    Genotype(edges=[(op, input_idx)] * 2*steps,
             steps=[StepGenotype(inner_edges=[(op, idx)] * 2*node_steps,
                                 inner_steps=[primitive name] * node_steps,
                                 inner_concat=[state idx ...])] * steps,
             concat=[state idx ...])
"""
from collections import namedtuple

Genotype = namedtuple('Genotype', 'edges steps concat')
StepGenotype = namedtuple('StepGenotype', 'inner_edges inner_steps inner_concat')

# candidate operations on a cell edge / an inner edge / a step node
PRIMITIVES = ['none', 'skip']
STEP_EDGE_PRIMITIVES = ['none', 'skip']
STEP_STEP_PRIMITIVES = ['Sum', 'ScaleDotAttn', 'LinearGLU', 'ConcatFC']

# names used by genotypes published before the primitives were renamed
# (reference genotypes.py:24-35 keeps this map in a comment)
LEGACY_STEP_NAMES = {'sum': 'Sum', 'scale_dot_attn': 'ScaleDotAttn',
                     'cat_conv_glu': 'LinearGLU', 'cat_conv_relu': 'ConcatFC'}


def upgrade_legacy(genotype):
    """Map a genotype that uses the pre-rename primitive names onto the current ones."""
    steps = [StepGenotype(inner_edges=list(s.inner_edges),
                          inner_steps=[LEGACY_STEP_NAMES.get(n, n) for n in s.inner_steps],
                          inner_concat=list(s.inner_concat)) for s in genotype.steps]
    return Genotype(edges=list(genotype.edges), steps=steps, concat=list(genotype.concat))
