"""The search step as one (or two) CUDA graphs, batch-sharded over the GPUs of a box.

One search step (train_searchable/ntu.py:70-93 + architect.py:21-29) is
    arch step   : fwd + bwd on a dev batch,   Adam(lr 3e-4, betas (0.5, .999), wd 1e-3) on alpha/beta/gamma
    weight step : fwd + bwd on a train batch, Adam(lr = cosine-restart schedule, wd) on the fusion weights
Both halves are captured once (after a warm-up that builds the launch plans) and replayed;
the learning rate is a device scalar written before each replay, inputs are copied into
static buffers.  Data parallelism replaces nn.DataParallel (ntu_darts_searchable.py:50-51):
one process per GPU, every rank holds a replica, the batch is sharded by sample, and ONE
NCCL all-reduce over the flat gradient arena [weights | alpha,beta,gamma | classifier]
sits between backward and the fused Adam (1/world folded into Adam's grad_scale).
BatchNorm uses per-replica batch statistics, as nn.DataParallel does.
"""
import torch

from . import native as N
from .optim import FusedAdam


class SearchStep:
    def __init__(self, head, criterion, B, num_classes, loss_kind='ce', eta_max=1e-3, eta_min=1e-6, Ti=1, Tm=2,
                 nbpe=100.0, weight_decay=3e-4, arch_lr=3e-4, arch_wd=1e-3, use_graphs=True, group=None,
                 full_fidelity=True):
        from models.auxiliary.scheduler import LRCosineAnnealingScheduler
        self.head, self.criterion = head, criterion
        self.device = next(head.parameters()).device
        self.B, self.kind = B, loss_kind
        a = head.args
        self.n_in, self.C, self.L = a.num_input_nodes, a.C, a.L
        self.group = group
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        self.use_graphs = use_graphs
        self.w_opt = FusedAdam(head.central_params(), lr=eta_max, weight_decay=weight_decay)
        self.a_opt = FusedAdam(head.arch_parameters(), lr=arch_lr, betas=(0.5, 0.999), weight_decay=arch_wd)
        self.w_opt.grad_scale = self.a_opt.grad_scale = 1.0 / self.world
        self.sched = LRCosineAnnealingScheduler(eta_max, eta_min, Ti, Tm, nbpe)
        dev = self.device
        # static input buffers: feats as views of one flat tensor so a batch arrives with ONE copy
        self.flat = {k: torch.zeros(self.n_in, B, self.C, self.L, device=dev) for k in ('dev', 'train')}
        self.feats = {k: [v[i] for i in range(self.n_in)] for k, v in self.flat.items()}
        if loss_kind == 'ce':
            self.labels = {k: torch.zeros(B, dtype=torch.int64, device=dev) for k in ('dev', 'train')}
        else:
            self.labels = {k: torch.zeros(B, num_classes, device=dev) for k in ('dev', 'train')}
        self.loss = {'dev': None, 'train': None}
        self.graphs = {}
        self.launches_per_step = None
        self.steps_done = 0
        # input pipeline (prefetch()): a copy stream fills the static buffers of the NEXT half step while the
        # current one computes; _ready[which] = copy finished, _done[which] = last reader of the buffers finished
        self.copy_stream = None
        self._ready = {'dev': None, 'train': None}
        self._done = {'dev': None, 'train': None}

    # ------------------------------------------------------------------ one half step, eager
    def _half(self, which):
        head = self.head
        loss = self.criterion(head(self.feats[which]), self.labels[which])
        loss.backward()
        self.allreduce_grads()
        (self.a_opt if which == 'dev' else self.w_opt).step()
        # detach: holding the autograd graph would keep the AccumulateGrad nodes (and the stream they were
        # created on) alive across steps, which breaks CUDA-graph capture on another stream
        return loss.detach()

    def allreduce_grads(self):
        """ONE collective per half step: sum the flat gradient arena [fusion weights | arch | classifier] over the
        ranks (the 1/world factor is applied inside the fused Adam as grad_scale)"""
        if self.world > 1:
            from .program import join_side
            join_side(self.device)
            torch.distributed.all_reduce(self.head._joint_arena(self.device).flat, group=self.group)

    def _run_half(self, which):
        ev = self._ready[which]
        if ev is not None:                   # a prefetch()ed batch: the compute stream waits for its copy
            torch.cuda.current_stream().wait_event(ev)
            self._ready[which] = None
        g = self.graphs.get(which)
        if g is not None:
            g.replay()
        else:
            self.loss[which] = self._half(which)
        if self.copy_stream is not None:     # the next prefetch() into these buffers must wait for this reader
            d = self._done[which]
            if d is None:
                d = self._done[which] = torch.cuda.Event()
            d.record(torch.cuda.current_stream())

    def load(self, which, feats_flat, labels):
        """copy a batch (any device, e.g. pinned host memory) into the static buffers; async"""
        self.flat[which].copy_(feats_flat, non_blocking=True)
        self.labels[which].copy_(labels, non_blocking=True)

    def prefetch(self, which, feats_flat, labels):
        """load() for an input pipeline: the copy runs on a dedicated stream, ordered after the last half step that
        read these buffers and before the next one that will, so the host->device transfer of the next batch
        overlaps the half step in flight (the 'dev' batch of step i+1 travels during the weight half of step i,
        the 'train' batch during the arch half of step i+1).  Source in pinned memory for a truly async copy."""
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        cs = self.copy_stream
        d = self._done[which]
        if d is not None:
            cs.wait_event(d)
        else:                                # no half step has run through the pipeline yet: order after everything queued
            cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            self.flat[which].copy_(feats_flat, non_blocking=True)
            self.labels[which].copy_(labels, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        self._ready[which] = ev

    # ------------------------------------------------------------------ state snapshot (warm-up must not train)
    def _snapshot(self):
        import copy
        snap = {'sd': {k: v.detach().clone() for k, v in self.head.state_dict().items()},
                'arch': [a.detach().clone() for a in self.head.arch_parameters()],
                'sched': copy.deepcopy(self.sched.__dict__), 'steps': self.steps_done}
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            sd = self.head.state_dict()
            for k, v in snap['sd'].items():
                sd[k].copy_(v)
            for a, b in zip(self.head.arch_parameters(), snap['arch']):
                a.copy_(b)
            for opt in (self.w_opt, self.a_opt):
                for st in opt._g.values():
                    st['m'].zero_()
                    st['v'].zero_()
                    st['step'].zero_()
        self.sched.__dict__.update(snap['sched'])
        self.steps_done = snap['steps']

    def prepare(self, warmup=3, restore=True):
        """warm-up (builds launch plans, arenas, Adam tables) then capture the two graphs; with restore=True
        the weights, architecture, BN buffers, Adam moments and schedule are put back afterwards"""
        self.head.train()
        snap = self._snapshot() if restore else None
        for _ in range(warmup):
            self.step()
        if not self.use_graphs:
            if restore:
                self._restore(snap)
            return
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.step()                      # one eager step on the capture stream (allocator warm-up)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = N.LAUNCHES[0]
        self.loss = {'dev': None, 'train': None}
        for which in ('dev', 'train'):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self.loss[which] = self._half(which)
            self.graphs[which] = g
        self.launches_per_step = N.LAUNCHES[0] - n0
        if restore:
            self._restore(snap)
        torch.cuda.synchronize()

    def step(self):
        """one search step on whatever currently sits in the static buffers; returns (arch loss, weight loss)
        as device scalars (no host sync)."""
        self._run_half('dev')
        lr = self.sched.step()
        self.w_opt.set_lr(float(self.sched.eta))
        self._run_half('train')
        self.steps_done += 1
        return self.loss['dev'], self.loss['train']

    def genotype(self):
        return self.head.genotype()

    def sync_buffers(self):
        """BatchNorm running statistics follow rank 0, as under nn.DataParallel"""
        if self.world > 1:
            for b in self.head.buffers():
                torch.distributed.broadcast(b, 0, group=self.group)
