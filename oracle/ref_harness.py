"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product (bm-nas_b200/).

Drives the UNMODIFIED reference modules vendored by tools/vendor_ref.sh into oracle/_ref/ (git-ignored copy of
/root/reference's hot-path files) through the reference's own loop body, so that bench.py's `--impl reference`
arm, `cpu_baseline` and the eager-GPU baseline time the real thing:

    model      = fusion head of Searchable_Skeleton_Image_Net (ntu_darts_searchable.py:94-101, 150):
                 FusionNetwork(steps, multiplier, num_input_nodes, 2, args, criterion) + nn.Linear
    optimizer  = Adam(params, lr=eta_max, weight_decay)                       (ntu_darts_searchable.py:39-42)
    scheduler  = LRCosineAnnealingScheduler(eta_max, eta_min, Ti, Tm, nbpe)   (:43-44)
    arch_opt   = Adam(arch_parameters, 3e-4, betas=(0.5, 0.999), wd=1e-3)     (:46-47)
    architect  = Architect(model, args, criterion, arch_opt)                  (:54)
    one step   = architect.step(dev batch) ; [no-grad metrics forward on the dev batch] ;
                 optimizer.zero_grad(); forward; scheduler.step(); scheduler.update_optimizer(); backward; step;
                 loss.item()                                                   (train_searchable/ntu.py:70-100)

Must run in its own process: the product package shadows the import path `models.search.darts` on purpose (it is
the drop-in), so this module puts oracle/_ref FIRST on sys.path and refuses to run if `models` is already imported
from somewhere else.
"""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available():
    return os.path.exists(os.path.join(REF, 'models', 'search', 'darts', 'model_search.py'))


def _import_reference():
    if not available():
        raise RuntimeError('oracle/_ref is empty: run tools/vendor_ref.sh where /root/reference exists')
    m = sys.modules.get('models')
    if m is not None and not os.path.abspath(getattr(m, '__file__', '') or '').startswith(REF):
        raise RuntimeError('a different `models` package is already imported; run the reference arm in its own process')
    for name, attrs in (('IPython', {'embed': lambda *a, **k: None}), ('graphviz', {'Digraph': object})):
        if name not in sys.modules:              # optional imports of the reference
            try:
                __import__(name)
            except Exception:
                mod = types.ModuleType(name)
                mod.__dict__.update(attrs)
                sys.modules[name] = mod
    sys.path.insert(0, REF)
    from models.search.darts.model_search import FusionNetwork       # import FIRST (circular import, SURVEY C-11)
    from models.search.darts.architect import Architect
    from models.auxiliary.scheduler import LRCosineAnnealingScheduler
    return FusionNetwork, Architect, LRCosineAnnealingScheduler


def build(c, device='cpu', seed=2):
    """c: a bench.py CONFIGS entry.  Returns a dict with the reference objects wired like the search script."""
    import torch
    import torch.nn as nn
    FusionNetwork, Architect, Sched = _import_reference()
    torch.manual_seed(seed)                                           # main_darts_searchable_ntu.py:17
    args = types.SimpleNamespace(C=c['C'], L=c['L'], drpt=c['drpt'], num_input_nodes=c['num_input_nodes'],
                                 steps=c['steps'], multiplier=c['multiplier'], node_steps=c['node_steps'],
                                 node_multiplier=c['node_multiplier'], weight_decay=c['weight_decay'], parallel=False)
    criterion = nn.CrossEntropyLoss() if c['loss'] == 'ce' else nn.BCEWithLogitsLoss()

    class Head(nn.Module):
        def __init__(self):
            super().__init__()
            self.fusion_net = FusionNetwork(steps=args.steps, multiplier=args.multiplier,
                                            num_input_nodes=args.num_input_nodes, num_keep_edges=2, args=args,
                                            criterion=criterion)
            self.central_classifier = nn.Linear(args.C * args.L * args.multiplier, c['classes'])

        def forward(self, feats):
            return self.central_classifier(self.fusion_net(feats))

        def arch_parameters(self):
            return self.fusion_net.arch_parameters()

    model = Head().to(device)       # nn.Module.to() moves weights only: alpha/beta/gamma stay on the CPU, as shipped
    model.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=c['eta_max'], weight_decay=c['weight_decay'])
    scheduler = Sched(c['eta_max'], 1e-6, 1, 2, 400.0)
    arch_opt = torch.optim.Adam(model.arch_parameters(), lr=3e-4, betas=(0.5, 0.999), weight_decay=1e-3)
    architect = Architect(model, args, criterion, arch_opt)
    return dict(model=model, optimizer=optimizer, scheduler=scheduler, architect=architect, criterion=criterion,
                args=args)


def search_step(R, dev, train, full_fidelity=False):
    """one search step of the reference loop body (see module docstring); returns the train loss (host float)"""
    import torch
    model, optimizer, scheduler = R['model'], R['optimizer'], R['scheduler']
    R['architect'].step(dev[0], dev[1], None)
    if full_fidelity:                     # dev phase: no-grad metrics forward, train-mode BN/dropout (ntu.py:77-85)
        optimizer.zero_grad()
        with torch.set_grad_enabled(False):
            out = model(dev[0])
            R['criterion'](out, dev[1]).item()
    optimizer.zero_grad()
    out = model(train[0])
    loss = R['criterion'](out, train[1])
    scheduler.step()
    scheduler.update_optimizer(optimizer)
    loss.backward()
    optimizer.step()
    return loss.item()


def time_search(c, steps, warmup, max_seconds, device='cpu', threads=None, full_fidelity=False):
    """(samples/s, ms/step, threads, steps done) of the reference search step on synthetic (B,C,L) features"""
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    if device != 'cpu':
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    R = build(c, device)
    g = torch.Generator().manual_seed(7)

    def batch():
        f = [torch.randn(c['B'], c['C'], c['L'], generator=g).to(device) for _ in range(c['num_input_nodes'])]
        if c['loss'] == 'ce':
            y = torch.randint(0, c['classes'], (c['B'],), generator=g).to(device)
        else:
            y = (torch.rand(c['B'], c['classes'], generator=g) < 0.2).float().to(device)
        return f, y
    pool = [batch() for _ in range(4)]

    def sync():
        if device != 'cpu':
            torch.cuda.synchronize()
    for i in range(warmup):
        search_step(R, pool[(2 * i) % 4], pool[(2 * i + 1) % 4], full_fidelity)
    sync()
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        search_step(R, pool[(2 * i) % 4], pool[(2 * i + 1) % 4], full_fidelity)
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    sync()
    dt = time.perf_counter() - t0
    return c['B'] * done / dt, 1e3 * dt / done, threads, done
