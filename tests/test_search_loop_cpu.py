"""CPU tests of the search-loop host glue: models/search/train_searchable/ntu.py and
models/search/darts/utils.py drop-ins.  The loop is device-agnostic host code, so it is exercised with a small
plain-torch stand-in for Searchable_Skeleton_Image_Net (same attribute names: reshape_layers, fusion_net,
genotype()).  When the reference checkout is present (build container) the same run is repeated through the
REFERENCE loop and everything observable must match: returned accuracy and genotype, log lines, saved files."""
import importlib.util
import logging
import os
import pickle
import sys
import types

import pytest
import torch
import torch.nn as nn

from helpers import ROOT

REF = '/root/reference'


class ToyNet(nn.Module):
    def __init__(self, Genotype):
        super().__init__()
        torch.manual_seed(0)
        self.reshape_layers = nn.ModuleList([nn.Linear(6, 5), nn.Linear(4, 5)])
        self.fusion_net = nn.Linear(10, 3)
        self.alpha = torch.zeros(2, requires_grad=True)
        self._G = Genotype

    def forward(self, feats):
        rgb, ske = feats
        h = torch.cat([self.reshape_layers[0](rgb), self.reshape_layers[1](ske)], 1)
        return self.fusion_net(torch.relu(h)) * (1 + self.alpha.sum())

    def genotype(self):
        k = int(self.alpha[0].item() > self.alpha[1].item())
        return self._G(edges=[('skip', k), ('skip', 1)], steps=[], concat=[2, 3])


class ToyArchitect:
    def __init__(self, model):
        self.model, self.calls = model, 0

    def step(self, input_valid, target_valid, logger):
        self.calls += 1
        with torch.no_grad():
            self.model.alpha += torch.tensor([0.01, -0.02]) * (1 if self.calls % 3 else -4)


def make_data(seed, n_batches, B):
    g = torch.Generator().manual_seed(seed)
    return [{'rgb': torch.randn(B, 6, generator=g), 'ske': torch.randn(B, 4, generator=g),
             'label': torch.randint(0, 3, (B,), generator=g)} for _ in range(n_batches)]


class ListLogger:
    def __init__(self):
        self.lines = []

    def info(self, s):
        self.lines.append(str(s))


class ListPlotter:
    def __init__(self):
        self.calls = []

    def plot(self, genotype, file_name):
        self.calls.append((str(genotype), os.path.basename(file_name)))


def run_loop(train_fn, sched_mod, Genotype, save_dir, status):
    os.makedirs(os.path.join(save_dir, 'best'), exist_ok=True)
    os.makedirs(os.path.join(save_dir, 'architectures'), exist_ok=True)
    model = ToyNet(Genotype)
    arch = ToyArchitect(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    sched = sched_mod.LRCosineAnnealingScheduler(1e-2, 1e-5, 1, 2, 3)
    loaders = {'train': make_data(1, 3, 8), 'dev': make_data(2, 2, 8), 'test': make_data(3, 2, 8)}
    sizes = {'train': 24, 'dev': 16, 'test': 16}
    logger, plotter = ListLogger(), ListPlotter()
    args = types.SimpleNamespace(save=save_dir)
    acc, geno = train_fn(model, arch, nn.CrossEntropyLoss(), opt, sched, loaders, sizes, device=torch.device('cpu'),
                         num_epochs=3, parallel=False, logger=logger, plotter=plotter, args=args, status=status)
    return dict(acc=float(acc), geno=str(geno), lines=logger.lines, plots=plotter.calls, arch_calls=arch.calls,
                sd={k: v.clone() for k, v in model.state_dict().items()}, files=sorted(os.listdir(os.path.join(save_dir, 'best'))))


def _ours():
    pkg = os.path.join(ROOT, 'bm-nas_b200')
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    import models.search.train_searchable.ntu as loop
    import models.auxiliary.scheduler as sched
    from models.search.darts.genotypes import Genotype
    return loop, sched, Genotype


@pytest.mark.parametrize('status', ['search', 'eval'])
def test_loop_bookkeeping_and_files(status, tmp_path):
    loop, sched, Genotype = _ours()
    r = run_loop(loop.train_ntu_track_acc, sched, Genotype, str(tmp_path), status)
    second = 'dev' if status == 'search' else 'test'
    assert r['arch_calls'] == (3 * 2 if status == 'search' else 0)          # Architect.step once per dev batch
    assert sum(l.startswith('train Loss') for l in r['lines']) == 3
    assert sum(l.startswith(second + ' Loss') for l in r['lines']) == 3
    assert [p[1] for p in r['plots']] == ['epoch_0', 'epoch_1', 'epoch_2']
    tag = 'best' if status == 'search' else 'best_test'
    assert f'{tag}_model.pt' in r['files'] and f'{tag}_genotype.pkl' in r['files']
    # the saved genotype is the reference's pickle format and the saved weights load back
    from models.search.darts.utils import load_pickle, load
    g = load_pickle(os.path.join(str(tmp_path), 'best', f'{tag}_genotype.pkl'))
    assert type(g).__module__ == 'models.search.darts.genotypes' and g._fields == ('edges', 'steps', 'concat')
    m2 = ToyNet(Genotype)
    load(m2, os.path.join(str(tmp_path), 'best', f'{tag}_model.pt'))
    assert 0.0 <= r['acc'] <= 1.0


def test_eval_helper_and_utils(tmp_path):
    loop, sched, Genotype = _ours()
    from models.search.darts import utils as U
    model = ToyNet(Genotype)
    logger = ListLogger()
    acc = loop.test_ntu_track_acc(model, {'test': make_data(3, 2, 8)}, nn.CrossEntropyLoss(), model.genotype(),
                                  {'test': 16}, torch.device('cpu'), logger, types.SimpleNamespace(save='x'))
    assert 0.0 <= float(acc) <= 1.0 and any(l.startswith('test Loss') for l in logger.lines)
    assert U.count_parameters(model) == sum(p.numel() for p in model.parameters())
    m = U.AvgrageMeter(); m.update(2.0, 2); m.update(4.0, 2)
    assert m.avg == 3.0
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1]])
    assert [float(a) for a in U.accuracy(out, torch.tensor([1, 2]), topk=(1, 2))] == [50.0, 50.0]
    U.create_exp_dir(str(tmp_path / 'exp'))
    assert sorted(os.listdir(str(tmp_path / 'exp'))) == ['architectures', 'best']


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (GPU box)')
@pytest.mark.parametrize('status', ['search', 'eval'])
def test_loop_matches_the_reference_loop(status, tmp_path):
    """the reference's own train_ntu_track_acc on the same toy model, data and seeds: identical accuracy, genotype,
    log lines, plots, saved files and final weights"""
    loop, sched, Genotype = _ours()
    ours = run_loop(loop.train_ntu_track_acc, sched, Genotype, str(tmp_path / 'ours'), status)
    # import the reference loop under private module names (its packages are also called ``models``)
    ip = types.ModuleType('IPython'); ip.embed = lambda *a, **k: None
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == 'models' or k.startswith('models.')}
    saved_path = list(sys.path)
    try:
        for k in saved:
            del sys.modules[k]
        sys.modules.setdefault('IPython', ip)
        sys.path.insert(0, REF)
        import models.search.train_searchable.ntu as ref_loop
        import models.auxiliary.scheduler as ref_sched
        from models.search.darts.genotypes import Genotype as RefGenotype
        ref = run_loop(ref_loop.train_ntu_track_acc, ref_sched, RefGenotype, str(tmp_path / 'ref'), status)
    finally:
        for k in [k for k in sys.modules if k == 'models' or k.startswith('models.')]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
        sys.path[:] = saved_path
    assert abs(ours['acc'] - ref['acc']) < 1e-12 and ours['geno'] == ref['geno']
    strip = lambda ls: [l for l in ls if not l.startswith('EXP:')]
    assert strip(ours['lines']) == strip(ref['lines'])
    assert ours['plots'] == ref['plots'] and ours['files'] == ref['files'] and ours['arch_calls'] == ref['arch_calls']
    for k in ours['sd']:
        assert torch.equal(ours['sd'][k], ref['sd'][k]), k


# ------------------------------------------------------------------ EgoGesture and MM-IMDB loops
class ToyEgoNet(ToyNet):
    """inputs arrive as (rgb (B,3,T,H,W), depth (B,1,T,H,W)) slices of one tensor"""

    def forward(self, feats):
        rgb, depth = feats
        return super().forward((rgb.flatten(1)[:, :6], depth.flatten(1)[:, :4]))


def make_ego_data(seed, n_batches, B):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(B, 4, 2, 2, 2, generator=g), torch.randint(0, 3, (B,), generator=g)) for _ in range(n_batches)]


class ToyMMNet(ToyNet):
    def forward(self, feats):
        text, image = feats
        return super().forward((image, text))


def make_mm_data(seed, n_batches, B):
    g = torch.Generator().manual_seed(seed)
    return [{'image': torch.randn(B, 6, generator=g), 'text': torch.randn(B, 4, generator=g),
             'label': (torch.rand(B, 3, generator=g) < 0.4).float()} for _ in range(n_batches)]


class TaskPlotter(ListPlotter):
    def plot(self, genotype, file_name, task=None):
        self.calls.append((str(genotype), os.path.basename(file_name), task))


class LRArchitect(ToyArchitect):
    def log_learning_rate(self, logger):
        logger.info("Architecture Learning Rate: {}".format(3e-4))


def run_task_loop(task, loop_mod, sched_mod, Genotype, save_dir, status, num_epochs=2):
    os.makedirs(os.path.join(save_dir, 'best'), exist_ok=True)
    os.makedirs(os.path.join(save_dir, 'architectures'), exist_ok=True)
    model = (ToyEgoNet if task == 'ego' else ToyMMNet)(Genotype)
    arch = LRArchitect(model)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    sched = sched_mod.LRCosineAnnealingScheduler(1e-2, 1e-5, 1, 2, 3)
    mk = make_ego_data if task == 'ego' else make_mm_data
    loaders = {'train': mk(1, 3, 8), 'dev': mk(2, 2, 8), 'test': mk(3, 2, 8)}
    sizes = {'train': 24, 'dev': 16, 'test': 16}
    logger, plotter = ListLogger(), TaskPlotter()
    args = types.SimpleNamespace(save=save_dir)
    if task == 'ego':
        res, geno = loop_mod.train_ego_track_acc(model, arch, nn.CrossEntropyLoss(), opt, sched, loaders, sizes,
                                                 device=torch.device('cpu'), num_epochs=num_epochs, parallel=False, logger=logger,
                                                 plotter=plotter, args=args, status=status)
    else:
        res, geno = loop_mod.train_mmimdb_track_f1(model, arch, nn.BCEWithLogitsLoss(), opt, sched, loaders, sizes, torch.device('cpu'),
                                                   num_epochs, False, logger, plotter, args, status=status)
    return dict(res=float(res), geno=str(geno), lines=logger.lines, plots=plotter.calls, arch_calls=arch.calls,
                sd={k: v.clone() for k, v in model.state_dict().items()}, files=sorted(os.listdir(os.path.join(save_dir, 'best'))))


def _our_loop(task):
    _ours()
    import importlib
    return importlib.import_module(f'models.search.train_searchable.{task}')


@pytest.mark.parametrize('task', ['ego', 'mmimdb'])
@pytest.mark.parametrize('status', ['search', 'eval'])
def test_ego_mmimdb_loop_bookkeeping(task, status, tmp_path):
    _, sched, Genotype = _ours()
    r = run_task_loop(task, _our_loop(task), sched, Genotype, str(tmp_path), status)
    assert r['arch_calls'] == (2 * 2 if status == 'search' else 0)
    assert [p[1:] for p in r['plots']] == [('epoch_0', task), ('epoch_1', task)]
    assert sum(l.startswith('Architecture Learning Rate') for l in r['lines']) == (2 if (task == 'mmimdb' or status == 'search') else 0)
    if task == 'mmimdb':
        assert sum('weighted F1' in l and l.startswith('dev Loss') for l in r['lines']) == 2
        assert 0.0 <= r['res'] <= 1.0
    else:
        assert sum(l.startswith('Learning Rate:') for l in r['lines']) == 4       # every phase (ego.py:53-55)
    tag = 'best' if status == 'search' else 'best_test'
    assert f'{tag}_genotype.pkl' in r['files']


def test_mmimdb_nan_failsafes(tmp_path):
    """a NaN training loss returns best_f1 right away with the model in eval mode (mmimdb.py:150-153)"""
    _, sched, Genotype = _ours()
    loop = _our_loop('mmimdb')
    os.makedirs(os.path.join(str(tmp_path), 'best'), exist_ok=True)
    model = ToyMMNet(Genotype)
    with torch.no_grad():
        model.fusion_net.weight.fill_(float('nan'))
    logger = ListLogger()
    out = loop.train_mmimdb_track_f1(model, LRArchitect(model), nn.BCEWithLogitsLoss(), torch.optim.Adam(model.parameters(), lr=1e-2),
                                     sched.LRCosineAnnealingScheduler(1e-2, 1e-5, 1, 2, 3),
                                     {'train': make_mm_data(1, 2, 8), 'dev': make_mm_data(2, 1, 8)}, {'train': 16, 'dev': 8},
                                     torch.device('cpu'), 1, False, logger, TaskPlotter(), types.SimpleNamespace(save=str(tmp_path)))
    assert out == 0.0 and not model.training and any('Nan loss during training' in l for l in logger.lines)


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (GPU box)')
@pytest.mark.parametrize('task', ['ego', 'mmimdb'])
@pytest.mark.parametrize('status', ['search', 'eval'])
def test_ego_mmimdb_loops_match_the_reference(task, status, tmp_path):
    """the reference's own loops on the same toy model, data and seeds: identical result, genotype, log lines, plots, files
    and final weights"""
    _, sched, Genotype = _ours()
    ours = run_task_loop(task, _our_loop(task), sched, Genotype, str(tmp_path / 'ours'), status)
    ip = types.ModuleType('IPython'); ip.embed = lambda *a, **k: None
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == 'models' or k.startswith('models.')}
    saved_path = list(sys.path)
    try:
        for k in saved:
            del sys.modules[k]
        sys.modules.setdefault('IPython', ip)
        sys.path.insert(0, REF)
        import importlib
        ref_loop = importlib.import_module(f'models.search.train_searchable.{task}')
        import models.auxiliary.scheduler as ref_sched
        from models.search.darts.genotypes import Genotype as RefGenotype
        ref = run_task_loop(task, ref_loop, ref_sched, RefGenotype, str(tmp_path / 'ref'), status)
    finally:
        for k in [k for k in sys.modules if k == 'models' or k.startswith('models.')]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
        sys.path[:] = saved_path
    assert abs(ours['res'] - ref['res']) < 1e-12 and ours['geno'] == ref['geno']
    strip = lambda ls: [l for l in ls if not l.startswith('EXP:')]
    assert strip(ours['lines']) == strip(ref['lines'])
    assert ours['plots'] == ref['plots'] and ours['files'] == ref['files'] and ours['arch_calls'] == ref['arch_calls']
    for k in ours['sd']:
        assert torch.equal(ours['sd'][k], ref['sd'][k]), k
