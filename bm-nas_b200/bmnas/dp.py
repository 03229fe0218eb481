"""Batch-sharded data parallelism over NVLink peer memory: the optimiser step fused with its collective.

Replaces nn.DataParallel (ntu_darts_searchable.py:50-51: scatter the batch, replicate the module, reduce-add the
gradients onto GPU 0, step there, re-broadcast next forward) for the search step.  One process per GPU; every rank holds a
replica whose parameters AND gradients live in symmetric (peer-mapped) memory; after a half step's backward every rank
launches ONE kernel, bmnas_dp_adam_step (csrc/dp_adam.cu):
    reduce-scatter of the half's gradient bucket (P2P loads) -> Adam on the rank's 1/world shard -> all-gather of the
    updated parameters (P2P stores into every replica),
i.e. ncclAllReduce + Adam-on-every-rank collapsed into one launch that moves each gradient element over NVLink once per
direction and does 1/world of the optimiser arithmetic per GPU.  Two buckets: [alpha, beta, gamma] for the Architect's
step (architect.py:21-24) and the weights for the weight step (train_searchable/ntu.py:88-93).
Falls back (PeerStep.create returns None) when symmetric memory cannot be set up; SearchStep then uses NCCL + FusedAdam.
"""
import ctypes

import torch

from . import native as N
from . import runtime as _rt


class _Bucket:
    pass


class PeerStep:
    @staticmethod
    def create(head, group, device, w_hyper, a_hyper, lr_ring=0):
        try:
            return PeerStep(head, group, device, w_hyper, a_hyper, lr_ring)
        except Exception as e:  # no symmetric memory on this platform / build: NCCL path
            import warnings
            warnings.warn(f'bmnas.dp: peer-memory optimiser step unavailable ({e!r}); using NCCL all-reduce + FusedAdam')
            return None

    def __init__(self, head, group, device, w_hyper, a_hyper, lr_ring=0):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        arch = list(head.arch_parameters())
        weights = [p for p in head.fusion_net.parameters() if p.requires_grad] + \
                  [p for p in head.central_classifier.parameters() if p.requires_grad]
        leaves = arch + weights
        pad4 = lambda n: (n + 3) // 4 * 4
        n_arch, n_all = sum(pad4(t.numel()) for t in arch), sum(pad4(t.numel()) for t in leaves)
        # shard boundaries fall on float4: every bucket length is a multiple of 4 by construction
        gflat = symm.empty(n_all, dtype=torch.float32, device=device)
        pflat = symm.empty(n_all, dtype=torch.float32, device=device)
        gflat.zero_()
        pflat.zero_()
        arena = _rt.GradArena(leaves, device, flat=gflat)
        head.__dict__['_bm_joint'] = arena
        head.fusion_net._bm_arena = arena
        head.central_classifier._bm_arena = arena
        head.fusion_net.__dict__.pop('_bm_cache', None)
        with torch.no_grad():
            for t in leaves:                      # parameters move into the symmetric buffer, same layout as the gradients
                o, cnt = arena.offsets[id(t)]
                view = pflat[o:o + cnt].view(t.shape)
                view.copy_(t.data)
                t.data = view
        self.gh = symm.rendezvous(gflat, group)
        self.ph = symm.rendezvous(pflat, group)
        self.gflat, self.pflat, self.arena = gflat, pflat, arena
        pad = self.gh.get_signal_pad(self.rank)
        pad.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)
        self.buckets = {}
        for which, lo, n, hyp, base in (('dev', 0, n_arch, a_hyper, 0), ('train', n_arch, n_all - n_arch, w_hyper, 2 * self.world)):
            b = _Bucket()
            b.lo, b.n = lo, n
            tbl = lambda ptrs, off: torch.tensor([int(p) + off for p in ptrs], dtype=torch.int64, device=device)
            b.gptrs, b.pptrs = tbl(self.gh.buffer_ptrs, lo * 4), tbl(self.ph.buffer_ptrs, lo * 4)
            b.sptrs = tbl(self.gh.signal_pad_ptrs, 0)
            per = ((n // 4 + self.world - 1) // self.world) * 4
            b.m, b.v = torch.zeros(per, device=device), torch.zeros(per, device=device)
            b.step = torch.zeros(1, dtype=torch.int64, device=device)
            b.lr_ring = lr_ring if which == 'train' else 0       # the weight step follows the schedule, the arch lr is constant
            b.lr = torch.full((max(1, b.lr_ring),), float(hyp['lr']), device=device)
            b.epoch = torch.zeros(1, dtype=torch.int32, device=device)
            b.done = torch.zeros(1, dtype=torch.int32, device=device)
            st = N.bmnas_dp_adam_params()
            st.world, st.rank, st.n = self.world, self.rank, n
            st.grad_ptrs, st.param_ptrs, st.signal_ptrs = b.gptrs.data_ptr(), b.pptrs.data_ptr(), b.sptrs.data_ptr()
            st.signal_base = base
            st.m, st.v, st.lr, st.step = b.m.data_ptr(), b.v.data_ptr(), b.lr.data_ptr(), b.step.data_ptr()
            st.beta1, st.beta2 = hyp['betas']
            st.eps, st.weight_decay, st.grad_scale = hyp.get('eps', 1e-8), hyp['weight_decay'], 1.0 / self.world
            st.epoch, st.done_counter = b.epoch.data_ptr(), b.done.data_ptr()
            st.lr_ring = b.lr_ring
            b.st = st
            self.buckets[which] = b

    def set_lr(self, lr):
        self.buckets['train'].lr.fill_(float(lr))

    def write_lr_ring(self, start, values):
        """FusedAdam.write_lr_ring for the weight bucket"""
        b = self.buckets['train']
        idx = torch.tensor([(start + i) % b.lr_ring for i in range(len(values))], dtype=torch.int64, device=self.device)
        b.lr.index_copy_(0, idx, torch.tensor(values, dtype=torch.float32).to(self.device))

    def step(self, which):
        """reduce-scatter + Adam + all-gather for this half's bucket: one launch, stream ordered, graph capturable"""
        from .program import join_side
        join_side(self.device)
        b = self.buckets[which]
        N.launch('bmnas_dp_adam_step', ctypes.byref(b.st), N.current_stream(self.device))
        _rt.clear_dirty(self.arena.tensors)

    def grad_span(self, which):
        b = self.buckets[which]
        return self.gflat[b.lo:b.lo + b.n]

    def state_snapshot(self):
        return {k: {f: getattr(b, f).clone() for f in ('m', 'v', 'step', 'lr')} for k, b in self.buckets.items()}

    def state_restore(self, snap):
        with torch.no_grad():
            for k, b in self.buckets.items():
                for f in ('m', 'v', 'step', 'lr'):
                    getattr(b, f).copy_(snap[k][f])
