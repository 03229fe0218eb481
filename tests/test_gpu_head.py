"""-m gpu: bmnas_head_fused (classifier + criterion + their backward in one launch, csrc/head_fused.cu) against plain
PyTorch on the CPU (fp32 as the reference, fp64 as the referee), through the C ABI and through SearchHead.loss_fused.
Tolerance: 1e-5 relative (north_star fp32 gate) on logits, loss, d loss / d logits and the input gradient."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from helpers import O, assert_close
import gpu_util as U

pytestmark = pytest.mark.gpu
TOL = 1e-5

SHAPES = [
    # B, K, N, kind                     what
    (96, 2048, 60, 0),                # NTU: 2 chunks per CTA, resident
    (32, 6144, 23, 1),                # MM-IMDB: BCE-with-logits, 6 chunks per CTA
    (96, 2048, 83, 0),                # EgoGesture: two (class, pair) rounds per thread
    (96, 16384, 83, 0),               # Ego-large: 16 chunks per CTA, streamed through two slots
    (5, 516, 7, 0),                   # ragged: partial last chunk, ranks without a chunk, odd batch
    (1, 128, 3, 1),                   # one sample, one chunk
    (200, 1024, 128, 0),              # more sample groups than one wave of clusters, N at the limit
    (77, 2048, 60, 1),                # odd batch, BCE
]


def _ref(x, W, b, target, kind, dtype):
    x = x.detach().to(dtype).clone().requires_grad_(True)
    W, b = W.to(dtype), b.to(dtype)
    logits = x @ W.t() + b
    logits.retain_grad()
    if kind == 0:
        loss = F.cross_entropy(logits, target)
    else:
        loss = F.binary_cross_entropy_with_logits(logits, target.to(dtype))
    loss.backward()
    return logits.detach(), loss.detach(), logits.grad, x.grad


def _run(x, W, b, target, kind, want_gx=True, with_bias=True):
    from bmnas import native as N
    B, K = x.shape
    Nc = W.shape[0]
    dev = U.DEV
    xd, Wd, bd = x.to(dev).contiguous(), W.to(dev).contiguous(), b.to(dev).contiguous()
    td = target.to(dev).contiguous()
    st = N.bmnas_head_params()
    st.B, st.K, st.N, st.kind = B, K, Nc, kind
    logits = torch.full((B, Nc), float('nan'), device=dev)
    gl = torch.full((B, Nc), float('nan'), device=dev)
    gx = torch.full((B, K), float('nan'), device=dev) if want_gx else None
    loss = torch.full((), float('nan'), device=dev)
    st.x, st.W, st.bias = xd.data_ptr(), Wd.data_ptr(), (bd.data_ptr() if with_bias else None)
    if kind == 0:
        st.labels = td.data_ptr()
    else:
        st.targets = td.data_ptr()
    st.logits, st.loss, st.glogits, st.gx = logits.data_ptr(), loss.data_ptr(), gl.data_ptr(), (gx.data_ptr() if want_gx else None)
    st.partials, st.counter = 16, 16
    assert N.lib().bmnas_head_supported(ctypes.byref(st)) == 1
    n = int(N.lib().bmnas_head_partials_size(ctypes.byref(st)))
    part = torch.zeros(n, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    st.partials, st.counter = part.data_ptr(), cnt.data_ptr()
    for _ in range(2):                  # twice: the ticket counter must come back to zero
        N.launch('bmnas_head_fused', ctypes.byref(st), N.current_stream())
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0
    return logits.cpu(), loss.cpu(), gl.cpu(), (gx.cpu() if want_gx else None)


@pytest.mark.parametrize('B,K,Nc,kind', SHAPES)
def test_head_fused_vs_torch(B, K, Nc, kind):
    g = torch.Generator().manual_seed(B + K + Nc)
    x = torch.randn(B, K, generator=g)
    W = torch.randn(Nc, K, generator=g) / K ** 0.5
    b = torch.randn(Nc, generator=g) * 0.1
    if kind == 0:
        target = torch.randint(0, Nc, (B,), generator=g)
    else:
        target = (torch.rand(B, Nc, generator=g) < 0.2).float()
    logits, loss, gl, gx = _run(x, W, b, target, kind)
    r32, r64 = _ref(x, W, b, target, kind, torch.float32), _ref(x, W, b, target, kind, torch.float64)
    for ours, a, c, what in zip((logits, loss, gl, gx), r32, r64, ('logits', 'loss', 'dlogits', 'gx')):
        # 1e-5 of the tensor's max against the fp64 referee, or as close to it as the CPU fp32 reference itself (x3)
        lim = max(TOL * c.abs().max().item(), 3.0 * (a.double() - c).abs().max().item())
        err = (ours.double() - c).abs().max().item()
        assert err <= lim, f'{what}: {err:.3e} > {lim:.3e}'


def test_head_forward_only_and_no_bias():
    g = torch.Generator().manual_seed(7)
    x, W, b = torch.randn(10, 256, generator=g), torch.randn(11, 256, generator=g) / 16, torch.zeros(11)
    target = torch.randint(0, 11, (10,), generator=g)
    logits, loss, gl, gx = _run(x, W, b, target, 0, want_gx=False, with_bias=False)
    assert gx is None
    r = _ref(x, W, b, target, 0, torch.float64)
    assert_close(logits, r[0], TOL, 'logits')
    assert_close(loss, r[1], TOL, 'loss')
    assert_close(gl, r[2], TOL, 'dlogits')


def test_head_bad_label_poisons_the_loss():
    g = torch.Generator().manual_seed(9)
    x, W, b = torch.randn(12, 128, generator=g), torch.randn(5, 128, generator=g) / 11, torch.zeros(5)
    target = torch.randint(0, 5, (12,), generator=g)
    target[3] = 5
    logits, loss, gl, gx = _run(x, W, b, target, 0)
    assert torch.isnan(loss)
    assert torch.isnan(gl[3]).all() and torch.isnan(gx[3]).all()
    ok = [i for i in range(12) if i != 3]
    assert torch.isfinite(gl[ok]).all() and torch.isfinite(gx[ok]).all()


@pytest.mark.parametrize('kind', ['ce', 'bce'])
def test_loss_fused_matches_the_five_launch_chain(kind):
    """SearchHead.loss_fused against criterion(head(feats), labels) on the same module: loss, logits and every gradient
    (fusion weights, alpha/beta/gamma, classifier) -- the two paths share everything but the head."""
    from bmnas import runtime as _rt
    from bmnas import nn as bnn
    cfg = O.Cfg(32, 8, 4, 2, 2, 2, 2, 0.0)
    B, ncls = 12, 9
    P = O.init_params(cfg, ncls, seed=3, prefix='cell')
    arch = O.init_arch(cfg, seed=3, scale=0.5)
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=2, loss=kind)
    crit = bnn.CrossEntropyLoss() if kind == 'ce' else bnn.BCEWithLogitsLoss()
    res = {}
    for fused in (True, False):
        head = U.build_head(cfg, ncls, P, arch)
        head.train()
        for m in head.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        fs = [f.to(U.DEV) for f in feats]
        with _rt.static_io():
            if fused:
                out = head.loss_fused(fs, labels.to(U.DEV), crit)
                assert out is not None
                loss, logits = out
            else:
                logits = head(fs)
                loss = crit(logits, labels.to(U.DEV))
            bnn.UNIT_LOSS_GRAD[0] = fused
            try:
                loss.backward()
            finally:
                bnn.UNIT_LOSS_GRAD[0] = False
        torch.cuda.synchronize()
        res[fused] = (loss.detach().cpu(), logits.detach().cpu(), U.grads_by_name(head),
                      [a.grad.detach().cpu().clone() for a in head.arch_parameters()])
    assert_close(res[True][0], res[False][0], TOL, 'loss')
    assert_close(res[True][1], res[False][1], TOL, 'logits')
    for k, v in res[False][2].items():
        assert (v is None) == (res[True][2][k] is None), k
        if v is not None:
            assert_close(res[True][2][k], v, 3e-5, 'grad ' + k, atol=1e-7)
    for i, (a, b) in enumerate(zip(res[True][3], res[False][3])):
        assert_close(a, b, 3e-5, f'garch{i}', atol=1e-8)


def test_loss_fused_scaled_backward():
    """a caller that scales the loss before backward (UNIT_LOSS_GRAD off): gradients scale with it"""
    from bmnas import runtime as _rt
    from bmnas import nn as bnn
    cfg = O.Cfg(32, 8, 4, 2, 2, 1, 1, 0.0)
    B, ncls = 6, 5
    P = O.init_params(cfg, ncls, seed=4, prefix='cell')
    arch = O.init_arch(cfg, seed=4, scale=0.5)
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=5)
    crit = bnn.CrossEntropyLoss()
    gs = []
    for scale in (1.0, 3.0):
        head = U.build_head(cfg, ncls, P, arch)
        head.train()
        for m in head.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        with _rt.static_io():
            loss, _ = head.loss_fused([f.to(U.DEV) for f in feats], labels.to(U.DEV), crit)
            (loss * scale).backward()
        torch.cuda.synchronize()
        gs.append(U.grads_by_name(head))
    for k, v in gs[0].items():
        if v is not None:
            assert_close(gs[1][k], 3.0 * v, 3e-5, 'grad ' + k, atol=1e-7)
