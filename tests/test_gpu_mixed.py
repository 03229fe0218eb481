"""-m gpu parity tests of the fused NodeMixedOp kernel bmnas_mixed_fwd (conv GEMM + BatchNorm statistics + grid barrier
+ Sum / ScaledDotAttn / LinearGLU / ConcatFC + softmax(gamma)-weighted sum in one tcgen05 launch) against the CPU
oracle's NodeMixedOp (node_operations.py:118-120) and against the two-kernel path it replaces, over
  * partial tiles (B*L not a multiple of 64), one resident wave, and batches beyond it (tiles recomputed in pass 2),
  * L in {4, 8, 16}, injected masks and in-kernel Philox, train and eval BatchNorm, no-grad forward (Z not written).
fp32 tolerance (north_star): 1e-5 relative."""
import types

import pytest
import torch

from helpers import O, assert_close, close_vs_referee
import gpu_util as U

pytestmark = pytest.mark.gpu
TOL, GTOL = 1e-5, 3e-5
C = 128


@pytest.fixture(autouse=True)
def _fused_everywhere():
    """the fused kernel is the default only from program.FUSED_MIXED_MIN_B samples on; these tests want it at every size"""
    from bmnas import program
    old = program.FUSED_MIXED_MIN_B
    program.FUSED_MIXED_MIN_B = 0
    yield
    program.FUSED_MIXED_MIN_B = old


OPS_RELU = ['Sum', 'ScaleDotAttn', 'LinearGLU', 'ConcatFC']
# ReLU has a knife edge: where the BatchNorm output is within rounding distance of 0, two correct fp32 forwards
# disagree on the gradient mask and ONE flipped element moves a whole sample of gx and a whole row of dW (seen at
# B=2500: sample 2210, FC row 16).  The beyond-one-wave cases therefore search over the smooth CatConvMish
# (node_operations.py:66-82, registered at run time exactly as the reference allows) instead of ConcatFC.
OPS_MISH = ['Sum', 'ScaleDotAttn', 'LinearGLU', 'CatConvMish']


def _mixed(L, drpt=0.2, seed=0, ops=OPS_RELU):
    from models.search.darts import node_operations as nops
    torch.manual_seed(seed)
    saved = list(nops.STEP_STEP_PRIMITIVES)
    nops.STEP_STEP_OPS.setdefault('CatConvMish', lambda C_, L_, a: nops.CatConvMish(C_, a))
    nops.STEP_STEP_PRIMITIVES[:] = ops
    try:
        mod = nops.NodeMixedOp(C, L, types.SimpleNamespace(C=C, L=L, drpt=drpt))
    finally:
        nops.STEP_STEP_PRIMITIVES[:] = saved
    with torch.no_grad():                         # non-trivial BatchNorm / LayerNorm affines and running statistics
        for n, p in mod.named_parameters():
            if n.endswith('bn.weight') or n.endswith('ln.weight'):
                p.add_(0.3 * torch.randn_like(p))
            if n.endswith('bn.bias') or n.endswith('ln.bias'):
                p.add_(0.2 * torch.randn_like(p))
        for n, b in mod.named_buffers():
            if n.endswith('running_mean'):
                b.add_(0.1 * torch.randn_like(b))
            if n.endswith('running_var'):
                b.mul_(1.0 + 0.3 * torch.rand_like(b))
    return mod


def _oracle(mod, x, w, go, masks, training, L, drpt, dtype=torch.float32, ops=OPS_RELU):
    P = {'mix.' + k: v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()
         for k, v in mod.state_dict().items()}
    P = {k: v.cpu() for k, v in P.items()}
    names = O.trainable_names(P)
    leaves = {k: P[k].clone().requires_grad_(True) for k in names}
    Pl = dict(P); Pl.update(leaves)
    xc = x.detach().cpu().to(dtype).requires_grad_(True)
    wc = w.detach().cpu().to(dtype).requires_grad_(True)
    cfg = O.Cfg(C, L, 2, 1, 1, 1, 1, drpt, step_ops=ops)
    mk = None if masks is None else {'mix.' + k: v.cpu() for k, v in masks.items()}
    out = O.node_mixed(xc, xc, wc, Pl, 'mix', mk, training, cfg)
    if go is not None:
        out.backward(go.cpu().to(dtype))
    return out.detach(), xc.grad, wc.grad, {k: v.grad for k, v in leaves.items()}, Pl


def _masks(mod, B, L, seed):
    g = torch.Generator().manual_seed(seed)
    m = {}
    for name, d in mod.named_modules():
        if isinstance(d, torch.nn.Dropout) and d.p > 0:
            m[name] = (torch.rand(B, C, L, generator=g) >= d.p).to(torch.uint8)
    return m


def _fused_launches(mod):
    names = []
    for r in mod._bm_cache.values():
        names += [c.name for c in r.prog.fwd]
    return names


@pytest.mark.parametrize('B,L', [(37, 8), (96, 8), (8, 8), (300, 8), (2500, 8), (5000, 8), (64, 4), (50, 16), (1200, 16)])
def test_fused_mixed_vs_oracle(B, L):
    from bmnas import program
    assert program.FUSED_MIXED != '0'
    _check_vs_oracle(B, L, 'bmnas_mixed_fwd')


@pytest.mark.parametrize('B,L', [(96, 8), (37, 8), (8, 8), (148, 8), (1, 8), (64, 4), (50, 16), (3, 16), (74, 16)])
def test_small_fused_mixed_vs_oracle(B, L, monkeypatch):
    """bmnas_mixed_small_fwd (csrc/mixed_small.cu): the FFMA variant that serves batches smaller than the machine --
    partial tiles, one sample, a grid that exactly fills the 148 SMs (B=148, L=8 and B=74, L=16), every L"""
    from bmnas import program
    monkeypatch.setattr(program, 'FUSED_MIXED', '0')
    monkeypatch.setattr(program, 'FUSED_MIXED_SMALL', '1')
    _check_vs_oracle(B, L, 'bmnas_mixed_small_fwd')


def _check_vs_oracle(B, L, expect):
    ops = OPS_MISH if B * L > 4096 else OPS_RELU
    mod = _mixed(L, ops=ops).to(U.DEV).train()
    sd0 = {k: v.clone() for k, v in mod.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, C, L, generator=g).to(U.DEV).requires_grad_(True)
    w = torch.softmax(torch.randn(4, generator=g), -1).to(U.DEV).requires_grad_(True)
    go = torch.randn(B, C, L, generator=g).to(U.DEV)
    masks = _masks(mod, B, L, 2)
    U.inject_masks(mod, masks)
    out = mod(x, x, w)
    out.backward(go)
    torch.cuda.synchronize()
    assert expect in _fused_launches(mod), 'the fused kernel did not take this shape'
    ref = _oracle(_restore(_mixed(L, ops=ops), sd0), x, w, go, masks, True, L, 0.2, ops=ops)
    ref64 = _oracle(_restore(_mixed(L, ops=ops), sd0), x, w, go, masks, True, L, 0.2, torch.float64, ops=ops)
    close_vs_referee(out, ref[0], ref64[0], TOL, 'out')
    close_vs_referee(x.grad, ref[1], ref64[1], GTOL, 'gx')
    close_vs_referee(w.grad, ref[2], ref64[2], GTOL, 'gw')
    for k, p in mod.named_parameters():
        close_vs_referee(p.grad, ref[3]['mix.' + k], ref64[3]['mix.' + k], GTOL, k, atol=2e-5 if k.endswith('conv.bias') else 1e-7)
    sd = mod.state_dict()
    for k, v in ref[4].items():
        if 'running' in k or 'num_batches' in k:
            assert_close(sd[k[4:]], v.detach(), 1e-5, k)
    # eval mode (running statistics) and the no-grad train-mode forward (Z is not written)
    mod.eval()
    with torch.no_grad():
        ev = mod(x.detach(), x.detach(), w.detach())
    ev_ref = _oracle(mod, x, w, None, None, False, L, 0.2, ops=ops)
    assert_close(ev, ev_ref[0], TOL, 'eval out')
    mod.train()
    sd1 = {k: v.clone() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        ng = mod(x.detach(), x.detach(), w.detach())
    ng_ref = _oracle(_restore(_mixed(L, ops=ops), sd1), x, w, None, masks, True, L, 0.2, ops=ops)
    assert_close(ng, ng_ref[0], TOL, 'no-grad train-mode out')


def _restore(mod, sd):
    mod.load_state_dict({k: v.clone() for k, v in sd.items()})
    return mod


@pytest.mark.parametrize('B,L,kernel', [(96, 8, 'bmnas_mixed_fwd'), (700, 8, 'bmnas_mixed_fwd'), (2500, 8, 'bmnas_mixed_fwd'),
                                        (96, 8, 'bmnas_mixed_small_fwd'), (50, 16, 'bmnas_mixed_small_fwd'), (41, 4, 'bmnas_mixed_small_fwd')])
def test_fused_equals_two_kernel_path_philox(B, L, kernel, monkeypatch):
    """same Philox seed and step: the fused kernels and bmnas_conv_fwd + bmnas_node_fwd must draw the same masks and
    agree to rounding (both write Z, mean, rstd; the backward kernels are shared)"""
    from bmnas import program, rng
    outs = []
    big = kernel == 'bmnas_mixed_fwd'
    for fused in ('1', '0'):
        program.FUSED_MIXED = fused if big else '0'
        monkeypatch.setattr(program, 'FUSED_MIXED_SMALL', '0' if big else fused)
        try:
            rng.manual_seed(1234)
            mod = _mixed(L, seed=3, ops=OPS_MISH if B * L > 4096 else OPS_RELU).to(U.DEV).train()
            g = torch.Generator().manual_seed(1)
            x = torch.randn(B, C, L, generator=g).to(U.DEV).requires_grad_(True)
            w = torch.softmax(torch.randn(4, generator=g), -1).to(U.DEV).requires_grad_(True)
            go = torch.randn(B, C, L, generator=g).to(U.DEV)
            out = mod(x, x, w)
            out.backward(go)
            torch.cuda.synchronize()
            names = _fused_launches(mod)
            assert (kernel in names) == (fused == '1'), names
            assert ('bmnas_conv_fwd' in names) == (fused == '0'), names
            outs.append((out.detach().clone(), x.grad.clone(), w.grad.clone(),
                         {k: p.grad.clone() for k, p in mod.named_parameters()},
                         {k: v.clone() for k, v in mod.state_dict().items()}))
        finally:
            program.FUSED_MIXED = 'auto'
    a, b = outs
    assert_close(a[0], b[0], 2e-6, 'out fused vs two-kernel')
    assert ((a[0] == 0) == (b[0] == 0)).all()
    assert_close(a[1], b[1], 1e-5, 'gx')
    assert_close(a[2], b[2], 1e-5, 'gw')
    for k in a[3]:
        if not k.endswith('conv.bias'):      # BN-fed conv biases: analytically zero gradient, pure rounding noise
            assert_close(a[3][k], b[3][k], 2e-5, k, atol=1e-7)
    for k in a[4]:
        assert_close(a[4][k], b[4][k], 1e-5, k)


@pytest.mark.parametrize('B,L', [(96, 8), (2500, 8), (300, 16)])
def test_fused_bf16_operand_mode(B, L):
    """gemm mode 3: the fused kernel stages bf16 operands (kind::f16, fp32 accumulation in TMEM, fp32 epilogue and
    fp32 Z / statistics); north_star's reduced-precision gate: 2e-2 relative against the fp32 oracle."""
    from bmnas import native as N
    lib = N.lib()
    lib.bmnas_set_gemm_mode(3)
    try:
        ops = OPS_MISH
        mod = _mixed(L, ops=ops).to(U.DEV).train()
        sd0 = {k: v.clone() for k, v in mod.state_dict().items()}
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, C, L, generator=g).to(U.DEV).requires_grad_(True)
        w = torch.softmax(torch.randn(4, generator=g), -1).to(U.DEV).requires_grad_(True)
        go = torch.randn(B, C, L, generator=g).to(U.DEV)
        masks = _masks(mod, B, L, 2)
        U.inject_masks(mod, masks)
        out = mod(x, x, w)
        out.backward(go)
        torch.cuda.synchronize()
        names = _fused_launches(mod)
        assert 'bmnas_mixed_fwd' in names
        call = [c for r in mod._bm_cache.values() for c in r.prog.fwd if c.name == 'bmnas_mixed_fwd'][0]
        assert call.args[0]._obj.wimg_fmt == 2, 'bf16 weight image expected in gemm mode 3'
        ref = _oracle(_restore(_mixed(L, ops=ops), sd0), x, w, go, masks, True, L, 0.2, ops=ops)
        assert_close(out, ref[0], 2e-2, 'out (bf16 operands)')
        assert (out.detach().cpu() - ref[0]).abs().max().item() > 1e-6, 'suspiciously exact: is the bf16 path really running?'
        assert_close(x.grad, ref[1], 2e-2, 'gx')
        assert_close(w.grad, ref[2], 2e-2, 'gw')
        for k, p in mod.named_parameters():
            if not k.endswith('conv.bias'):
                assert_close(p.grad, ref[3]['mix.' + k], 2e-2, k)
    finally:
        lib.bmnas_set_gemm_mode(1)
