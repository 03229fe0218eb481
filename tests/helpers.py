"""Shared test helpers: golden-fixture loading and comparison utilities."""
import io
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bmnas_oracle as O  # noqa: E402


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return {k: z[k] for k in z.files}


def sub(d, prefix):
    """entries of d under 'prefix/' with the prefix stripped, as torch tensors."""
    n = len(prefix)
    return {k[n:]: torch.from_numpy(np.array(v)) for k, v in d.items() if k.startswith(prefix)}


def cfg_of(d, step_ops=None):
    g = lambda k: d['cfg_' + k].item()
    return O.Cfg(C=int(g('C')), L=int(g('L')), num_input_nodes=int(g('num_input_nodes')), steps=int(g('steps')),
                 multiplier=int(g('multiplier')), node_steps=int(g('node_steps')),
                 node_multiplier=int(g('node_multiplier')), drpt=float(g('drpt')),
                 step_ops=step_ops or O.STEP_STEP_PRIMITIVES)


def arch_of(d, prefix):
    a = sub(d, prefix)
    return [a[str(i)] for i in range(len(a))]


class _RefUnpickler(pickle.Unpickler):
    """Reference pickles name ``models.search.darts.genotypes``; map onto the oracle's
    namedtuples so CPU tests do not need the product package."""

    def find_class(self, module, name):
        if module == 'models.search.darts.genotypes':
            return getattr(O, name)
        return super().find_class(module, name)


def unpickle_genotype(arr):
    return _RefUnpickler(io.BytesIO(np.asarray(arr).tobytes())).load()


def geno_plain(g):
    """namedtuple-agnostic structural form of a genotype."""
    return (list(map(tuple, g.edges)),
            [(list(map(tuple, s.inner_edges)), list(s.inner_steps), list(s.inner_concat)) for s in g.steps],
            list(g.concat))


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return ((a - b).abs().max() / denom).item()


def assert_close(a, b, tol, what='', atol=0.0):
    """max|a-b| <= tol * max|b| + atol"""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    err = (a - b).abs().max().item() if a.numel() else 0.0
    lim = tol * (b.abs().max().item() if b.numel() else 0.0) + atol
    assert err <= lim, f'{what}: max abs err {err:.3e} > {lim:.3e} (rel {rel_err(a, b):.3e})'


def close_vs_referee(ours, ref32, ref64, tol, what, atol=0.0, knife=0, cpu_mult=3.0):
    """ours must be as close to the fp64 referee as tol, or as the fp32 CPU reference itself (x3):
    a gradient that is the difference of large terms is not computable to 1e-5 in fp32 by anyone.
    knife > 0 (large-batch gradient checks only): up to `knife` dim-0 slices (samples of an input gradient, output
    rows of a weight gradient) may miss the tolerance, by at most 10 % of the tensor's max.  ReLU is discontinuous in
    its gradient: where a BatchNorm output lies within fp32 rounding of 0, two correct forwards disagree on the mask
    and that ONE element moves its whole sample of gx and its whole row of dW (measured: B=2500, sample 2210, FC row
    16 -- tools/diag_mixed2.py); with millions of activations per tensor such an element exists with probability
    O(0.2) per test.  Every other slice still has to meet `tol`.
    cpu_mult: how many times the CPU-fp32 reference's own distance to the referee is accepted (3 by default; 10 for
    plans whose reductions run over >= 8192 columns or through >= 12 stacked mixed ops, where the fp32 accumulation ORDER
    -- tensor-core K-chunk order and split-K atomics here, MKL's blocked summation there -- sets the error)."""
    ours = torch.as_tensor(ours).double().cpu()
    r32, r64 = ref32.double(), ref64.double()
    lim = max(tol * r64.abs().max().item(), cpu_mult * (r32 - r64).abs().max().item()) + atol
    d = (ours - r64).abs()
    err = d.max().item() if d.numel() else 0.0
    if err <= lim:
        return
    if knife > 0 and d.dim() >= 1 and d.shape[0] > 1:
        per = d.reshape(d.shape[0], -1).max(dim=1).values
        bad = per > lim
        loose = 0.1 * r64.abs().max().item() + atol
        assert int(bad.sum()) <= knife and per.max().item() <= loose, (
            f'{what}: {int(bad.sum())} dim-0 slices miss {lim:.3e} (allowed {knife}), worst {per.max().item():.3e} (loose bound {loose:.3e})')
        return
    assert err <= lim, f'{what}: max abs err {err:.3e} > {lim:.3e}'
