"""The search step as one (or two) CUDA graphs, batch-sharded over the GPUs of a box.

One search step (train_searchable/ntu.py:70-93 + architect.py:21-29) is
    arch step   : fwd + bwd on a dev batch,   Adam(lr 3e-4, betas (0.5, .999), wd 1e-3) on alpha/beta/gamma
    weight step : fwd + bwd on a train batch, Adam(lr = cosine-restart schedule, wd) on the fusion weights
Both halves are captured once (after a warm-up that builds the launch plans) and replayed;
the learning rate is a device scalar written before each replay, inputs are copied into
static buffers.  Data parallelism replaces nn.DataParallel (ntu_darts_searchable.py:50-51):
one process per GPU, every rank holds a replica, the batch is sharded by sample, and ONE
NCCL all-reduce per half step sits between backward and the fused Adam (1/world folded into
Adam's grad_scale): the arch half reduces only the alpha/beta/gamma span of the flat gradient
arena (70 floats at NTU), the weight half only the weight span.
BatchNorm uses per-replica batch statistics, as nn.DataParallel does.

Discarded work is not done (prune_grads=True): the Architect's backward computes weight
gradients that train_searchable/ntu.py:77 zeroes unread, and the weight step computes
architecture gradients that architect.py:22 zeroes unread; the two launch plans are built
with runtime.grad_mode('arch') / ('weights') so neither set is produced.  The updates are
bit-identical to the reference's (tests/test_gpu_parity.py::test_search_loop_golden).

Schedule note: the reference loop runs a whole 'train' epoch of weight steps and then a whole
'dev' epoch of Architect steps (train_searchable/ntu.py:31-38); SearchStep.step() is ONE arch
half followed by ONE weight half (the unit SURVEY 8d defines as a step).  Callers that want the
reference's epoch schedule call half('train') / half('dev') themselves
(models/search/train_searchable/*.py do).
"""
import torch

from . import native as N
from .optim import FusedAdam

LR_RING = 1024      # device ring of weight-step learning rates (bmnas_adam_params.lr_ring); refilled half a ring at a time


class _LossHandle:
    """result of SearchStep.read_loss_async(): get() -> (arch loss, weight loss) as Python floats"""

    def __init__(self, host, ev):
        self.host, self.ev = host, ev

    def get(self):
        self.ev.synchronize()
        return float(self.host[0]), float(self.host[1])


class SearchStep:
    def __init__(self, head, criterion, B, num_classes, loss_kind='ce', eta_max=1e-3, eta_min=1e-6, Ti=1, Tm=2,
                 nbpe=100.0, weight_decay=3e-4, arch_lr=3e-4, arch_wd=1e-3, use_graphs=True, group=None,
                 prune_grads=True, sync_replicas=True, peer_step=None):
        from models.auxiliary.scheduler import LRCosineAnnealingScheduler
        self.head, self.criterion = head, criterion
        self.device = next(head.parameters()).device
        self.B, self.kind = B, loss_kind
        a = head.args
        self.n_in, self.C, self.L = a.num_input_nodes, a.C, a.L
        self.group = group
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        self.use_graphs = use_graphs
        self.prune_grads = prune_grads
        self.rank = torch.distributed.get_rank(group) if group is not None else 0
        if self.world > 1:
            from . import runtime as _rt
            # world-size-invariant dropout streams: Philox is keyed by the GLOBAL sample index
            _rt.SAMPLE_OFFSET[0] = self.rank * B
            if sync_replicas:
                self.sync_replicas()
        # world > 1: the optimiser step fused with its collective over NVLink peer memory (bmnas.dp); peer_step=False
        # forces the NCCL all-reduce + FusedAdam path, None tries the fused one and falls back
        self.peer = None
        if self.world > 1 and peer_step is not False and str(self.device).startswith('cuda') and not N.VALIDATE_ONLY:
            from .dp import PeerStep
            self.peer = PeerStep.create(head, group, self.device,
                                        dict(lr=eta_max, betas=(0.9, 0.999), weight_decay=weight_decay),
                                        dict(lr=arch_lr, betas=(0.5, 0.999), weight_decay=arch_wd), lr_ring=LR_RING)
        # the reference hands Adam two parameter groups with identical hyper-parameters (central_params(),
        # ntu_darts_searchable.py:38-42): one group here = one multi-tensor launch per weight step
        self.w_opt = FusedAdam([p for g in head.central_params() for p in g['params']], lr=eta_max, weight_decay=weight_decay)
        # the schedule is uploaded LR_RING / 2 steps ahead (FusedAdam.write_lr_ring): no host write between replays
        self.w_opt.lr_ring = LR_RING
        self.w_steps = 0            # host mirror of the weight optimiser's device step counter
        self._ring_until = -1       # weight steps < this have their learning rate on the device
        self.a_opt = FusedAdam(head.arch_parameters(), lr=arch_lr, betas=(0.5, 0.999), weight_decay=arch_wd)
        self.w_opt.grad_scale = self.a_opt.grad_scale = 1.0 / self.world
        self.sched = LRCosineAnnealingScheduler(eta_max, eta_min, Ti, Tm, nbpe)
        dev = self.device
        # static input buffers: ONE byte buffer per search step, [dev feats | train feats | dev labels | train labels]
        # (pack_step() builds that layout), so a whole step's inputs arrive with one copy (load_step) and a half's
        # features with one (load / prefetch); feats / labels are typed views into it
        nf = self.n_in * B * self.C * self.L * 4
        nl = B * 8 if loss_kind == 'ce' else B * num_classes * 4
        nl16 = (nl + 15) // 16 * 16
        self._layout = {'dev': (0, 2 * nf), 'train': (nf, 2 * nf + nl16), 'nf': nf, 'nl': nl, 'total': 2 * nf + 2 * nl16}
        self.packed = torch.zeros(self._layout['total'], dtype=torch.uint8, device=dev)
        self.flat, self.labels = {}, {}
        for k in ('dev', 'train'):
            fo, lo = self._layout[k]
            self.flat[k] = self.packed[fo:fo + nf].view(torch.float32).view(self.n_in, B, self.C, self.L)
            lab = self.packed[lo:lo + nl]
            self.labels[k] = lab.view(torch.int64) if loss_kind == 'ce' else lab.view(torch.float32).view(B, num_classes)
        self.feats = {k: [v[i] for i in range(self.n_in)] for k, v in self.flat.items()}
        self._one = torch.ones((), device=dev)
        self.loss = {'dev': None, 'train': None}
        self.logits = {'dev': None, 'train': None}     # logits of the last half step (fused head only; static buffers under graphs)
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
        from . import runtime as _rt
        from . import nn as _nn
        from . import program as _prog
        head = self.head
        mode = ('arch' if which == 'dev' else 'weights') if self.prune_grads else 'all'
        with _rt.grad_mode(mode), _rt.static_io(), _prog.early_zero():
            # classifier + criterion + the gradient they send back: one launch where the head takes the case
            res = head.loss_fused(self.feats[which], self.labels[which], self.criterion) if hasattr(head, 'loss_fused') else None
            loss = res[0] if res is not None else self.criterion(head(self.feats[which]), self.labels[which])
            self.logits[which] = res[1] if res is not None else None
            _nn.UNIT_LOSS_GRAD[0] = True       # the backward below is seeded with the constant 1 (a resident tensor: no fill launch)
            try:
                torch.autograd.backward(loss, grad_tensors=self._one)
            finally:
                _nn.UNIT_LOSS_GRAD[0] = False
        if self.peer is not None:
            self.peer.step(which)            # reduce-scatter + Adam on the shard + all-gather of parameters: one launch
        else:
            self.allreduce_grads(which)
            (self.a_opt if which == 'dev' else self.w_opt).step()
        # detach: holding the autograd graph would keep the AccumulateGrad nodes (and the stream they were
        # created on) alive across steps, which breaks CUDA-graph capture on another stream
        return loss.detach()

    def set_lr(self, lr):
        """a constant learning rate for the weight step (until the next half('train'), which follows the schedule)"""
        self.w_opt.set_lr(lr)
        if self.peer is not None:
            self.peer.set_lr(lr)
        self._ring_until = -1

    def _refill_lr(self):
        """upload the schedule values of the next LR_RING / 2 weight steps (scheduler.py:25-40, one tick per step)"""
        import copy
        c = copy.deepcopy(self.sched)
        vals = []
        for _ in range(LR_RING // 2):
            c.step()
            vals.append(float(c.eta))
        if self.peer is not None:
            self.peer.write_lr_ring(self.w_steps, vals)
            ok = True
        else:
            ok = self.w_opt.write_lr_ring(self.w_steps, vals)
        # before the optimiser's device state exists only the value of the very next step is in place (group['lr'])
        self._ring_until = self.w_steps + (len(vals) if ok else 1)

    def grad_span(self, which):
        """the contiguous slice of the flat gradient arena [alpha,beta,gamma | fusion weights | classifier] that the
        optimiser of this half step reads"""
        ar = self.head._joint_arena(self.device)
        if which == 'dev':
            return ar.span(self.head.arch_parameters())
        return ar.span([p for p in self.head.parameters() if p.requires_grad])

    def allreduce_grads(self, which=None):
        """ONE collective per half step: sum this half's span of the gradient arena over the ranks (the 1/world
        factor is applied inside the fused Adam as grad_scale).  which=None: the whole arena."""
        if self.world > 1:
            from .program import join_side
            join_side(self.device)
            buf = self.head._joint_arena(self.device).flat if which is None else self.grad_span(which)
            torch.distributed.all_reduce(buf, group=self.group)

    def sync_replicas(self):
        """rank 0's weights, BatchNorm buffers, architecture tensors and Philox seed become every rank's (differently
        seeded replicas would otherwise train apart silently: every rank applies the same reduced gradient)"""
        if self.world <= 1:
            return
        from . import rng
        with torch.no_grad():
            ts = list(self.head.parameters()) + list(self.head.buffers()) + list(self.head.arch_parameters())
            for t in ts:
                torch.distributed.broadcast(t.data, 0, group=self.group)
            sd = torch.tensor([rng._state['seed'], rng._state['n']], dtype=torch.int64, device=self.device)
            torch.distributed.broadcast(sd, 0, group=self.group)
            rng._state['seed'], rng._state['n'] = int(sd[0]), int(sd[1])

    def _run_half(self, which):
        ev = self._ready[which]
        if ev is not None:                   # a prefetch()ed batch: the compute stream waits for its copy
            torch.cuda.current_stream().wait_event(ev)
            self._ready[which] = None
        g = self.graphs.get(which)
        if g is not None:
            g.replay()
            self.loss[which], self.logits[which] = self._half_out[which]
        else:
            self.loss[which] = self._half(which)
        if which == 'train':
            self.w_steps += 1
        if self.copy_stream is not None:     # the next prefetch() into these buffers must wait for this reader
            d = self._done[which]
            if d is None:
                d = self._done[which] = torch.cuda.Event()
            d.record(torch.cuda.current_stream())

    def load(self, which, feats_flat, labels):
        """copy a batch (any device, e.g. pinned host memory) into the static buffers; async"""
        self.flat[which].copy_(feats_flat, non_blocking=True)
        self.labels[which].copy_(labels, non_blocking=True)

    def pack_step(self, dev_feats, dev_labels, train_feats, train_labels, pinned=False, device=None):
        """one contiguous byte tensor holding a whole step's inputs in the layout of the static buffers (for load_step);
        feats: (n_in, B, C, L) fp32, labels as the criterion takes them"""
        out = torch.empty(self._layout['total'], dtype=torch.uint8, device=device or 'cpu')
        if pinned:
            out = out.pin_memory()
        nf, nl = self._layout['nf'], self._layout['nl']
        for k, f, y in (('dev', dev_feats, dev_labels), ('train', train_feats, train_labels)):
            fo, lo = self._layout[k]
            out[fo:fo + nf].view(torch.float32).view(self.flat[k].shape).copy_(f)
            lab = out[lo:lo + nl]
            (lab.view(torch.int64) if self.kind == 'ce' else lab.view(torch.float32).view(self.labels[k].shape)).copy_(y)
        return out

    def load_step(self, packed):
        """both halves' inputs with ONE copy (packed: pack_step(); any device, e.g. pinned host memory); async"""
        if (packed.is_cuda and packed.device == self.packed.device and packed.dtype == torch.uint8 and packed.is_contiguous()
                and packed.numel() == self.packed.numel() and packed.data_ptr() % 16 == 0 and packed.numel() % 16 == 0):
            import ctypes
            N.launch('bmnas_copy', ctypes.c_void_p(self.packed.data_ptr()), ctypes.c_void_p(packed.data_ptr()),
                     ctypes.c_longlong(packed.numel()), N.current_stream(self.device))
        else:
            self.packed.copy_(packed, non_blocking=True)

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

    def prefetch_step(self, packed_host):
        """load_step() for an input pipeline: the packed inputs of the NEXT step (pack_step(..., pinned=True)) travel
        host -> device on a dedicated copy stream into one of two staging buffers while the current step computes; the
        next step() then moves them into the static buffers with one device copy (bmnas_copy, ~4 us) and replays the
        one-graph step.  The compute stream never waits for PCIe."""
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        if not hasattr(self, '_stage'):
            self._stage = [torch.empty_like(self.packed) for _ in range(2)]
            self._stage_free = [None, None]        # event: the device copy that last read staging buffer j has finished
            self._stage_n = 0
            self._staged = []                       # FIFO of (buffer index, 'copy landed' event)
        j = self._stage_n % 2
        self._stage_n += 1
        cs = self.copy_stream
        if self._stage_free[j] is not None:
            cs.wait_event(self._stage_free[j])
        with torch.cuda.stream(cs):
            self._stage[j].copy_(packed_host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        self._staged.append((j, ev))

    def reset_pipeline(self):
        """forget the input-pipeline state (after a synchronize): the next step() takes the one-graph path again"""
        torch.cuda.synchronize(self.device)
        self._ready = {'dev': None, 'train': None}
        self._done = {'dev': None, 'train': None}
        if hasattr(self, '_staged'):
            self._staged = []
            self._stage_free = [None, None]

    def _consume_staged(self):
        """the oldest prefetch_step() batch becomes the content of the static buffers (compute stream)"""
        j, ev = self._staged.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        self.load_step(self._stage[j])
        done = self._stage_free[j]
        if done is None:
            done = self._stage_free[j] = torch.cuda.Event()
        done.record(cur)

    def read_loss_async(self):
        """device -> host read of the last step's (arch loss, weight loss) without stalling the launch of the next step:
        the two scalars are copied into pinned host memory behind the step on the compute stream; .get() on the returned
        handle waits for that copy only.  A training loop calls it every step and consumes the value one step later."""
        if not hasattr(self, '_loss_ring'):
            self._loss_ring = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(4)]
            self._loss_dev = torch.empty(2, dtype=torch.float32, device=self.device)
            self._loss_n = 0
        host = self._loss_ring[self._loss_n % 4]
        self._loss_n += 1
        torch.stack([self.loss['dev'].reshape(()), self.loss['train'].reshape(())], out=self._loss_dev)
        host.copy_(self._loss_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return _LossHandle(host, ev)

    # ------------------------------------------------------------------ state snapshot (warm-up must not train)
    def _snapshot(self):
        import copy
        snap = {'sd': {k: v.detach().clone() for k, v in self.head.state_dict().items()},
                'arch': [a.detach().clone() for a in self.head.arch_parameters()],
                'sched': copy.deepcopy(self.sched.__dict__), 'steps': self.steps_done, 'w_steps': self.w_steps,
                'opt': [opt.state_snapshot() for opt in (self.w_opt, self.a_opt)],
                'peer': self.peer.state_snapshot() if self.peer is not None else None,
                'rng': [(p.rng_state.clone() if p.rng_state is not None else None) for p in self._programs()]}
        return snap

    def _hoist_pair(self):
        """(arch-half plan, weight-half plan) of the fusion network when the weight half's preamble may run during the arch
        half: pruned plans (two distinct programs), the gradient-span clear forked during the forward (the side branch the
        preamble rides on), a single device"""
        import os
        from . import program as _prog
        if not self.prune_grads or os.environ.get('BMNAS_HOIST_PREP', '1') == '0' or not _prog.SIDE_WGRAD:
            return None
        cache = self.head.fusion_net.__dict__.get('_bm_cache', {})
        pa = [r.prog for k, r in cache.items() if 'arch' in k and r.prog.training and r.prog.want_backward and r.prog.B == self.B]
        pw = [r.prog for k, r in cache.items() if 'weights' in k and r.prog.training and r.prog.want_backward and r.prog.B == self.B]
        if len(pa) != 1 or len(pw) != 1 or not pa[0]._zero_ranges or not pw[0]._prep_calls:
            return None
        return pa[0], pw[0]

    def _programs(self):
        return [r.prog for r in self.head.fusion_net.__dict__.get('_bm_cache', {}).values()]

    def _restore(self, snap):
        with torch.no_grad():
            sd = self.head.state_dict()
            for k, v in snap['sd'].items():
                sd[k].copy_(v)
            for a, b in zip(self.head.arch_parameters(), snap['arch']):
                a.copy_(b)
            for opt, osnap in zip((self.w_opt, self.a_opt), snap['opt']):
                opt.state_restore(osnap)          # moments / step counters as they were (zero if they did not exist)
            if self.peer is not None:
                self.peer.state_restore(snap['peer'])
            progs = self._programs()
            for p, r in zip(progs, snap['rng']):   # plans that existed at snapshot time get their step counter back;
                if r is not None and p.rng_state is not None:
                    p.rng_state.copy_(r)
            for p in progs[len(snap['rng']):]:     # plans built during the warm-up restart at step 0
                if p.rng_state is not None:
                    p.rng_state[1] = 0
        self.sched.__dict__.update(snap['sched'])
        self.steps_done = snap['steps']
        self.w_steps = snap['w_steps']
        self._ring_until = -1

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
        self._half_out = {}
        for which in ('dev', 'train'):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self.loss[which] = self._half(which)
            self.graphs[which] = g
            self._half_out[which] = (self.loss[which], self.logits[which])     # static outputs of this graph
        self.launches_per_step = N.LAUNCHES[0] - n0
        # the whole step as ONE graph as well: with the schedule on the device (lr ring) nothing happens on the host
        # between the two halves, so step() is a single launch and the arch half's Adam -> weight half's first kernel
        # boundary is an ordinary graph edge
        g = torch.cuda.CUDAGraph()
        self._step_out = {}
        hoist = self._hoist_pair()
        with torch.cuda.graph(g, stream=s):
            if hoist is not None:
                # the weight half's preamble (bmnas_wprep + Philox step) on the side branch of the arch half: the Architect
                # only moves alpha/beta/gamma, so the weights the images are made of are already final
                pa, pw = hoist

                def extra(sp, side, pw=pw):
                    pw.run_prep(sp)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    pw._prep_hoisted = ev
                pa.side_extra = extra
            try:
                for which in ('dev', 'train'):
                    loss = self._half(which)
                    self._step_out[which] = (loss, self.logits[which])
            finally:
                if hoist is not None:
                    hoist[0].side_extra = None
                    hoist[1]._prep_hoisted = None
        self.graphs['step'] = g
        if restore:
            self._restore(snap)
        torch.cuda.synchronize()

    def half(self, which):
        """one half step ('dev' = Architect.step, 'train' = weight step incl. the LR schedule tick) on the batch in
        the static buffers; returns the loss as a device scalar"""
        if which == 'train':
            if self.w_steps >= self._ring_until:
                self._refill_lr()
            self.sched.step()
            self.w_opt.note_lr(float(self.sched.eta))
        self._run_half(which)
        return self.loss[which]

    def step(self):
        """one search step on whatever currently sits in the static buffers; returns (arch loss, weight loss)
        as device scalars (no host sync)."""
        if getattr(self, '_staged', None):
            self._consume_staged()
        g = self.graphs.get('step')
        half_pipeline = self._ready['dev'] is not None or self._ready['train'] is not None or self._done['dev'] is not None
        if g is not None and not half_pipeline:                 # (the per-half prefetch() pipeline hands over per half step)
            if self.w_steps >= self._ring_until:
                self._refill_lr()
            self.sched.step()
            self.w_opt.note_lr(float(self.sched.eta))
            g.replay()
            self.w_steps += 1
            self.steps_done += 1
            for which in ('dev', 'train'):
                self.loss[which], self.logits[which] = self._step_out[which]
            return self.loss['dev'], self.loss['train']
        self.half('dev')
        self.half('train')
        self.steps_done += 1
        return self.loss['dev'], self.loss['train']

    def metrics_forward(self, which='dev'):
        """the reference's dev-phase metrics pass (train_searchable/ntu.py:81-85): a no-grad forward in train mode
        (batch statistics, fresh dropout masks, BN running statistics updated -- C-13) on the same batch the
        Architect just used.  Returns (loss, logits) as device tensors."""
        g = self.graphs.get('metrics:' + which)
        if g is not None:
            g.replay()
            return self._metrics[which]
        from . import runtime as _rt
        with torch.no_grad(), _rt.static_io():
            res = self.head.loss_fused(self.feats[which], self.labels[which], self.criterion) if hasattr(self.head, 'loss_fused') else None
            if res is not None:
                return res
            logits = self.head(self.feats[which])
            loss = self.criterion(logits, self.labels[which])
        return loss, logits

    def capture_metrics_forward(self, which='dev'):
        """capture metrics_forward as a CUDA graph (after prepare())"""
        if not self.use_graphs:
            return
        self.metrics_forward(which)
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.metrics_forward(which)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        if not hasattr(self, '_metrics'):
            self._metrics = {}
        with torch.cuda.graph(g, stream=s):
            self._metrics[which] = self.metrics_forward(which)
        self.graphs['metrics:' + which] = g

    def genotype(self):
        """alpha/beta/gamma are bit-identical on every rank (same reduced gradient, same Adam), so any rank may
        derive the genotype without a collective"""
        return self.head.genotype()

    def checkpoint(self):
        """what the reference saves (best_model.pt = state_dict, train_searchable/ntu.py:141-144) plus what it loses
        (SURVEY C-3): architecture tensors, both optimisers incl. the fused moments, the schedule.  BatchNorm running
        statistics are rank 0's, as under nn.DataParallel -- call on rank 0 (or sync_buffers() on all ranks first)."""
        return {'state_dict': {k: v.detach().cpu().clone() for k, v in self.head.state_dict().items()},
                'arch': [a.detach().cpu().clone() for a in self.head.arch_parameters()],
                'w_opt': self.w_opt.state_dict(), 'a_opt': self.a_opt.state_dict(),
                'sched': dict(self.sched.__dict__), 'steps': self.steps_done, 'w_steps': self.w_steps}

    def load_checkpoint(self, ck):
        with torch.no_grad():
            self.head.load_state_dict(ck['state_dict'])
            for a, b in zip(self.head.arch_parameters(), ck['arch']):
                a.copy_(b.to(a.device))
        self.w_opt.load_state_dict(ck['w_opt'])
        self.a_opt.load_state_dict(ck['a_opt'])
        self.sched.__dict__.update(ck['sched'])
        self.steps_done = ck['steps']
        self.w_steps = ck.get('w_steps', ck['steps'])
        self._ring_until = -1

    def sync_buffers(self):
        """BatchNorm running statistics follow rank 0, as under nn.DataParallel"""
        if self.world > 1:
            for b in self.head.buffers():
                torch.distributed.broadcast(b, 0, group=self.group)
