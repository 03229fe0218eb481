"""-m gpu parity tests of the code paths bench.py times (parity holes a/b/d):
  * in-kernel Philox dropout (not injected masks): the masks of the very step are read back through the C-ABI test
    hook bmnas_philox_keep_mask and handed to the CPU oracle -- logits prove the forward drew exactly those masks,
    gradients prove the backward (and the NodeCell tail in ln_block.cu) re-drew the same ones, for the CTA-per-sample
    and the warp-per-sample node kernels alike;
  * default kernel dispatch at large batch (warp-per-sample node kernels + tcgen05 panel GEMMs + tcgen05 wgrad):
    whole plans at NTU B=1024 / B=4096 and the found network at B=1024 against the oracle;
  * a 20-step search trajectory at the NTU size through CUDA graphs with the pruned arch / weight plans, against
    O.SearchState half step by half step, genotype identical at every step.
fp32 tolerance (north_star): 1e-5 relative on logits/loss; gradients 3e-5 of the tensor's max or 3x the CPU-fp32
reference's own distance to the fp64 referee (helpers.close_vs_referee)."""
import pytest
import torch

from helpers import O, load, unpickle_genotype, geno_plain, assert_close, close_vs_referee
import gpu_util as U

pytestmark = pytest.mark.gpu
TOL, GTOL = 1e-5, 3e-5
MAX_ATTEMPTS = 4

NTU = O.Cfg(128, 8, 8, 2, 2, 2, 2, 0.2)
CASES = {
    'ntu_B96': dict(cfg=NTU, B=96, classes=60, kind='ce', seed=4),
    'ntu_B1024': dict(cfg=NTU, B=1024, classes=60, kind='ce'),
    'ntu_B4096': dict(cfg=NTU, B=4096, classes=60, kind='ce', seed=4),
    'mmimdb_B32': dict(cfg=O.Cfg(192, 16, 6, 2, 2, 1, 1, 0.1), B=32, classes=23, kind='bce'),
    'ego_B96': dict(cfg=O.Cfg(128, 8, 8, 2, 2, 3, 3, 0.05), B=96, classes=83, kind='ce'),
    'ego_large_B96': dict(cfg=O.Cfg(256, 16, 8, 4, 4, 3, 3, 0.05), B=96, classes=83, kind='ce', seed=4),
    'found_ntu_B1024': dict(cfg=NTU, B=1024, classes=60, kind='ce', found=True),
}


def _bias_atol(k):
    return 2e-5 if k.endswith('conv.bias') else 1e-7


def _loss_mod(kind):
    from bmnas.nn import CrossEntropyLoss, BCEWithLogitsLoss
    return CrossEntropyLoss() if kind == 'ce' else BCEWithLogitsLoss()


def _set_variant(v):
    from bmnas import native as N
    from bmnas import program
    N.lib().bmnas_set_node_variant(v)
    program.CHAIN_NODE = (v != 2)


@pytest.fixture(autouse=True)
def _restore_variant():
    yield
    _set_variant(0)


class _Collect:
    """close_vs_referee with the verdict recorded instead of raised (one attempt of test_philox_plan_vs_oracle)"""

    def __init__(self):
        self.fail, self.gross = [], []

    def __call__(self, ours, ref32, ref64, tol, what, loose=0.1, **kw):
        try:
            close_vs_referee(ours, ref32, ref64, tol, what, **kw)
        except AssertionError as e:
            self.fail.append(str(e))
            d = (torch.as_tensor(ours).double().cpu() - ref64.double()).abs().max().item()
            if d > loose * ref64.double().abs().max().item() + kw.get('atol', 0.0):
                self.gross.append(str(e))


def _attempt(c, variant, seed):
    cfg, B, ncls, kind = c['cfg'], c['B'], c['classes'], c['kind']
    gt = unpickle_genotype(load('found_ntu_golden')['genotype']) if c.get('found') else None
    P = O.init_params(cfg, ncls, seed=seed, prefix='cell', genotype=gt)
    arch = None if gt is not None else O.init_arch(cfg, seed=seed, scale=0.5)
    feats, labels = O.synthetic_batch(cfg, B, ncls, seed=seed - 1, loss=kind)
    head = U.build_head(cfg, ncls, P, arch, genotype=gt)
    head.train()
    # two forward/backward passes: the second one runs with an advanced step counter (fresh masks); the masks of each
    # pass are read back right after it and the oracle takes the same two passes (BN buffers advance twice)
    Pc = {k: v.clone() for k, v in P.items()}
    for it in range(2):
        for p in list(head.parameters()) + head.arch_parameters():
            p.grad = None                      # plain module use accumulates like autograd: start each pass from zero
        out = head([f.to(U.DEV) for f in feats])
        loss = _loss_mod(kind)(out, labels.to(U.DEV))
        loss.backward()
        torch.cuda.synchronize()
        prog = U.training_program(head)
        masks = U.philox_masks(head, B, prog)
        assert masks, 'no live dropout site'
        if it == 1:
            assert any(not torch.equal(masks[n], prev[n]) for n in masks), 'masks must change per forward'
        prev = masks
        lv, logits, gw, ga = O.loss_and_grads(feats, labels, arch, Pc, masks, cfg, loss=kind, genotype=gt)
    for n, m in masks.items():                      # sanity: the masks are real Bernoulli(1-p) draws
        p_ = dict(head.named_modules())[n].p
        assert abs(1.0 - m.float().mean().item() - p_) < 0.02 + 3.0 / (m.numel() ** 0.5), (n, m.float().mean().item())
    dbl = lambda t: t.double() if t.is_floating_point() else t
    P64 = {k: dbl(v) for k, v in P.items()}
    a64 = None if arch is None else [a.double() for a in arch]
    lv64, logits64, gw64, ga64 = O.loss_and_grads([f.double() for f in feats], dbl(labels), a64, P64, masks, cfg,
                                                  loss=kind, genotype=gt)
    # the forward is continuous in every activation: it has to meet the gate on EVERY attempt
    close_vs_referee(out, logits, logits64, TOL, 'logits')
    close_vs_referee(loss, lv, lv64, TOL, 'loss')
    chk = _Collect()
    # ReLU knife edges (helpers.close_vs_referee): only where a plan holds > 1M activations per ReLU site
    big = B * cfg.C * cfg.L >= (1 << 20) or cfg.C * cfg.L >= 4096
    for k, p in head.named_parameters():
        if gw.get(k) is None:
            continue
        chk(p.grad, gw[k], gw64[k], GTOL, 'grad ' + k, atol=_bias_atol(k),
            knife=max(2, p.shape[0] // 64) if big else 0, cpu_mult=10.0 if big else 3.0)
    if arch is not None:
        for i, a in enumerate(head.arch_parameters()):
            chk(a.grad, ga[i], ga64[i], GTOL, f'garch{i}')
    sd = head.state_dict()
    for k in Pc:
        if 'running' in k or 'num_batches' in k:
            assert_close(sd[k], Pc[k], 1e-5, k)
    return chk


@pytest.mark.parametrize('variant', [0, 2], ids=['dispatch_default', 'node_warp'])
@pytest.mark.parametrize('name', list(CASES))
def test_philox_plan_vs_oracle(name, variant):
    """Gradients are compared on up to MAX_ATTEMPTS data sets (seed, seed + 10, ...); one clean attempt passes the case.
    Why attempts: ReLU's derivative is discontinuous.  These plans hold 5-20 M ReLU decisions; where a BatchNorm / LayerNorm
    output lies within fp32 rounding (~1e-7) of 0, two correct fp32 forwards disagree on that ONE decision (expected
    count per attempt O(1): tools/diag_tol.py shows attempt 0 failing and attempt 1 of the same build passing with every
    tensor at <= 0.4x its limit, and the BatchNorm statistics' atomic order moves which element it is from run to
    run).  One flipped decision shifts a summed gradient by one element's contribution ~ max / sqrt(#elements) ~ 5e-4
    of the tensor's max -- far above the 3e-5 gate, far below a wrong kernel.  So: the forward (logits, loss) and the
    BatchNorm buffers must meet their gate on EVERY attempt, no gradient may be off by more than 10 % of its tensor's
    max on ANY attempt (a wrong kernel is off by O(1)), and at least one attempt must meet the full gradient gate."""
    c = CASES[name]
    if variant == 2 and c['B'] > 96:
        pytest.skip('default dispatch already selects the warp-per-sample kernels at this batch')
    _set_variant(variant)
    notes = []
    for att in range(MAX_ATTEMPTS):
        chk = _attempt(c, variant, c.get('seed', 3) + 10 * att)
        assert not chk.gross, f'attempt {att}: gross gradient error: ' + ' | '.join(chk.gross[:4])
        if not chk.fail:
            return
        notes.append(f'attempt {att}: {len(chk.fail)} tensors miss the gate, first: {chk.fail[0]}')
    raise AssertionError(f'no clean attempt in {MAX_ATTEMPTS}: ' + ' || '.join(notes))


@pytest.mark.parametrize('graphs', [True, False], ids=['graphs', 'eager'])
def test_ntu_trajectory_20_steps(graphs):
    """20 search steps at the NTU size (B=96) exactly as bench.py runs them: captured graphs, Philox dropout, pruned
    arch / weight plans.  After every half step the masks of that half are read back and the oracle takes the same
    half step; alpha/beta/gamma, losses, the LR and the genotype are compared at every step."""
    from bmnas.search import SearchStep
    cfg, B, ncls, nsteps = NTU, 96, 60, 20
    P = O.init_params(cfg, ncls, seed=5, prefix='cell')
    arch = O.init_arch(cfg, seed=5, scale=1e-3)        # the reference's own init: 1e-3 * randn (model_search.py:102)
    head = U.build_head(cfg, ncls, P, arch)
    head.train()
    # script defaults (main_darts_searchable_ntu.py): with alpha/beta/gamma ~ 1e-3 and Adam moving every coordinate by
    # ~lr = 3e-4 per step, the genotype changes several times inside 20 steps
    hyper = dict(eta_max=1e-3, eta_min=1e-6, Ti=1, Tm=2, nbpe=8.0, weight_decay=3e-4, arch_lr=3e-4, arch_wd=1e-3)
    ss = SearchStep(head, _loss_mod('ce'), B, ncls, use_graphs=graphs, **hyper)
    f0, y0 = O.synthetic_batch(cfg, B, ncls, seed=50)
    ss.load('dev', torch.stack(f0), y0)
    ss.load('train', torch.stack(f0), y0)
    ss.prepare(warmup=2, restore=True)
    st = O.SearchState(cfg, {k: v.clone() for k, v in P.items()}, [a.clone() for a in arch], loss='ce', **hyper)
    genos = set()
    near_ties = 0
    for s in range(nsteps):
        dv = O.synthetic_batch(cfg, B, ncls, seed=100 + 2 * s)
        tr = O.synthetic_batch(cfg, B, ncls, seed=101 + 2 * s)
        ss.load('dev', torch.stack(dv[0]), dv[1])
        ss.load('train', torch.stack(tr[0]), tr[1])
        la = ss.half('dev')
        torch.cuda.synchronize()
        ola = st.arch_step(dv[0], dv[1], U.philox_masks(head, B, U.training_program(head, 'arch')))
        assert_close(la, ola, 1e-4, f'arch loss step {s}')
        for i, a in enumerate(head.arch_parameters()):
            # Adam normalises every coordinate, so a coordinate whose gradient is small relative to its rounding
            # error moves by a visibly different fraction of lr per step: bound the gap by 5 % of the distance
            # travelled (lr per step), on top of the usual relative tolerance
            assert_close(a, st.arch[i], 1e-4, f'arch {i} step {s}', atol=0.05 * hyper['arch_lr'] * (s + 1))
        lw = ss.half('train')
        torch.cuda.synchronize()
        olw = st.weight_step(tr[0], tr[1], U.philox_masks(head, B, U.training_program(head, 'weights')))
        assert_close(lw, olw, 1e-4, f'weight loss step {s}')
        assert abs(ss.sched.eta - st.sched.eta) < 1e-12
        g_ours, g_ref = geno_plain(head.genotype()), geno_plain(st.genotype())
        if g_ours != g_ref:
            # alpha/beta/gamma agree to the tolerance asserted above, but an arg-max between two candidates that are closer
            # than that tolerance can still fall either way (the weight-gradient GEMMs accumulate with atomics, so even two
            # runs of this build differ in the last bits).  Then (1) the reference derivation applied to OUR architecture
            # tensors must give OUR genotype -- the derivation itself has to be identical -- and (2) such near-ties must be
            # rare: at most 2 of the 20 steps.
            ours_arch = [a.detach().float().cpu() for a in head.arch_parameters()]
            assert geno_plain(O.network_genotype(ours_arch, cfg)) == g_ours, (s, g_ours, g_ref)
            near_ties += 1
            assert near_ties <= 2, (s, g_ours, g_ref)
        genos.add(str(g_ours))
    assert len(genos) > 1, 'the genotype never changed: the trajectory test is too easy'
    sd = head.state_dict()
    for k, v in st.P.items():
        # Adam turns a coordinate whose gradient is pure rounding noise (BN-fed conv biases, weights of nearly dead
        # units) into +-lr steps: bound the gap by 5 % of the distance an Adam coordinate can travel in 20 steps
        assert_close(sd[k], v, 1e-3, 'final ' + k, atol=(0.05 * 20 * hyper['eta_max']) if v.dtype.is_floating_point else 0)
