"""-m gpu tests of the autograd contract of the drop-in modules: outside SearchStep the modules behave like the
reference nn.Modules -- gradients accumulate across backward passes until zero_grad()/step(), outputs and input
gradients are fresh tensors -- while SearchStep keeps its static-buffer, overwrite-per-backward fast path."""
import pytest
import torch

from helpers import O, assert_close
import gpu_util as U

pytestmark = pytest.mark.gpu


def _head():
    cfg = O.Cfg(32, 8, 4, 2, 2, 2, 2, 0.0)
    P = O.init_params(cfg, 5, seed=1, prefix='cell')
    arch = O.init_arch(cfg, seed=1, scale=0.3)
    head = U.build_head(cfg, 5, P, arch)
    head.train()
    for m in head.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return cfg, head


def test_gradients_accumulate_until_zero_grad():
    from bmnas.nn import CrossEntropyLoss
    from bmnas.optim import FusedAdam
    cfg, head = _head()
    crit = CrossEntropyLoss()
    b1 = O.synthetic_batch(cfg, 6, 5, seed=2)
    b2 = O.synthetic_batch(cfg, 6, 5, seed=3)
    f1, f2 = [f.to(U.DEV) for f in b1[0]], [f.to(U.DEV) for f in b2[0]]

    def grads_of(batches):
        for p in list(head.parameters()) + head.arch_parameters():
            p.grad = None
        from bmnas import runtime as rt
        rt.clear_dirty(list(head.parameters()) + head.arch_parameters())
        for f, y in batches:
            crit(head(f), y.to(U.DEV)).backward()
        torch.cuda.synchronize()
        return [p.grad.detach().clone() for p in list(head.parameters()) + head.arch_parameters()]

    # running statistics move with every forward; BatchNorm uses batch statistics in train mode, so gradients do not
    g1 = grads_of([(f1, b1[1])])
    g2 = grads_of([(f2, b2[1])])
    g12 = grads_of([(f1, b1[1]), (f2, b2[1])])             # two backward passes, no zero_grad in between
    for a, b, c in zip(g1, g2, g12):
        assert_close(c, a + b, 1e-5, 'accumulated gradient', atol=1e-7)
    # an optimiser step (or zero_grad) ends the accumulation window
    opt = FusedAdam(head.central_params(), lr=1e-3)
    opt.zero_grad()
    crit(head(f1), b1[1].to(U.DEV)).backward()
    torch.cuda.synchronize()
    g1b = [p.grad.detach().clone() for p in head.parameters()]
    for a, b in zip(g1[:len(g1b)], g1b):
        assert_close(b, a, 1e-5, 'gradient after zero_grad', atol=1e-7)


def test_outputs_and_input_grads_are_fresh_tensors():
    from bmnas.nn import CrossEntropyLoss
    cfg, head = _head()
    b1 = O.synthetic_batch(cfg, 6, 5, seed=2)
    b2 = O.synthetic_batch(cfg, 6, 5, seed=3)
    x1 = [f.to(U.DEV).requires_grad_(True) for f in b1[0]]
    x2 = [f.to(U.DEV).requires_grad_(True) for f in b2[0]]
    o1 = head(x1)
    o1_copy = o1.detach().clone()
    o2 = head(x2)
    assert o1.data_ptr() != o2.data_ptr()
    assert torch.equal(o1.detach(), o1_copy), 'a second forward overwrote the first output'
    crit = CrossEntropyLoss()
    crit(o2, b2[1].to(U.DEV)).backward()
    g2 = [x.grad.clone() for x in x2]
    o1b = head(x1)
    crit(o1b, b1[1].to(U.DEV)).backward()
    for a, x in zip(g2, x2):
        assert torch.equal(a, x.grad), 'a later backward overwrote an earlier input gradient'


def test_out_of_range_label_poisons_the_loss():
    from bmnas.nn import CrossEntropyLoss
    logits = torch.randn(8, 5, device=U.DEV, requires_grad=True)
    y = torch.tensor([0, 1, 2, 3, 4, -100, 1, 2], device=U.DEV)
    loss = CrossEntropyLoss()(logits, y)
    loss.backward()
    assert torch.isnan(loss).item() and torch.isnan(logits.grad[5]).all() and torch.isfinite(logits.grad[0]).all()


def test_second_order_architect_vs_oracle():
    """the unrolled DARTS update (the reference names it with --unrolled, main_darts_found_ntu.py:48, but
    ships only the first-order Architect) against the oracle's restatement of the same finite-difference formula"""
    import types
    from bmnas.nn import CrossEntropyLoss
    from models.search.darts.architect import Architect
    cfg, head = _head()
    P0 = {k: v.detach().cpu().clone() for k, v in head.state_dict().items()}
    arch0 = [a.detach().cpu().clone() for a in head.arch_parameters()]
    tr = O.synthetic_batch(cfg, 6, 5, seed=2)
    va = O.synthetic_batch(cfg, 6, 5, seed=3)
    eta, wd = 0.05, 3e-4

    class Capture(torch.optim.Optimizer):            # records the gradient Architect hands to the optimiser
        def __init__(self, params):
            super().__init__(params, {})
            self.seen = None

        def step(self):
            self.seen = [p.grad.detach().cpu().clone() for g in self.param_groups for p in g['params']]

    opt = Capture(head.arch_parameters())
    arc = Architect(head, types.SimpleNamespace(weight_decay=wd), CrossEntropyLoss(), opt)
    buf0 = [b.detach().clone() for b in head.buffers()]
    w0 = [p.detach().clone() for p in head.parameters()]
    arc.step_unrolled([f.to(U.DEV) for f in tr[0]], tr[1].to(U.DEV), [f.to(U.DEV) for f in va[0]], va[1].to(U.DEV), eta)
    torch.cuda.synchronize()
    want = O.unrolled_arch_grad(tr, va, arch0, P0, cfg, eta, wd)
    for got, ref in zip(opt.seen, want):
        # a finite difference of two fp32 gradients divided by 2 eps: the quotient amplifies rounding by eta / (2 eps)
        assert_close(got, ref, 2e-3, 'unrolled architecture gradient', atol=1e-6)
    for a, b in zip(head.parameters(), w0):
        assert torch.equal(a.detach(), b), 'weights must be restored after the virtual steps'
    for a, b in zip(head.buffers(), buf0):
        assert torch.equal(a, b), 'BatchNorm buffers must be restored'
