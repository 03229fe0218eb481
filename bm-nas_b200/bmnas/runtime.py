"""Glue between torch.autograd and the static launch plans (program.py).

* GradArena  -- one flat fp32 gradient buffer per root module; every parameter's and
               architecture tensor's ``.grad`` is a view into it.  The backward kernels
               write into the views directly (one bucket for the NCCL all-reduce, one
               descriptor table for the fused Adam).  Semantics: each backward
               OVERWRITES the arena (as if ``zero_grad(set_to_none=True)`` preceded it),
               which is what the reference loop does (train_searchable/ntu.py:77,
               architect.py:22).
* Runner     -- binds per-call pointers, runs a Program forward/backward, hands the
               gradients to autograd.
* run()      -- cache lookup + autograd.Function entry used by every drop-in module.
"""
import torch

from . import native as N
from .program import Program, Slot


import weakref

# Plain module use (model(x); loss.backward(); several losses / micro-batches before one optimizer.step()) gets autograd's
# semantics: gradients ACCUMULATE across backward passes until zero_grad()/step(), and outputs / input gradients are fresh
# tensors.  SearchStep (and anything else that owns the whole step) switches to STATIC_IO: outputs and input gradients
# are views of the plan's static buffers and every backward overwrites the arena -- what the captured graphs need and what
# the reference loop's zero_grad-before-every-backward amounts to (train_searchable/ntu.py:77, architect.py:22).
STATIC_IO = [False]
_ARENAS = weakref.WeakSet()


class static_io:
    def __enter__(self):
        self.prev = STATIC_IO[0]
        STATIC_IO[0] = True

    def __exit__(self, *a):
        STATIC_IO[0] = self.prev


def clear_dirty(tensors):
    """optimizer.step() / zero_grad(): the next backward into these leaves starts from zero again"""
    ids = {id(t) for t in tensors}
    for ar in list(_ARENAS):
        ar.dirty -= ids


SAMPLE_OFFSET = [0]      # global index of this rank's first sample (world-size-invariant dropout streams)

# Which leaf gradients the next launch plans produce.  'all' (default, plain autograd use of the modules): every
# parameter and architecture tensor.  'arch': alpha/beta/gamma only -- the Architect's step (architect.py:21-29)
# back-propagates through everything but steps only the architecture tensors and the loop zeroes the weight
# gradients unread (train_searchable/ntu.py:77), so the plan omits every weight-gradient GEMM, the classifier dW
# and the BN/LN affine gradients.  'weights': parameters only -- the weight step's architecture gradients are
# likewise cleared unread by the next arch_optimizer.zero_grad() (architect.py:22).  The reference distinguishes the
# two sets the same way: architecture tensors are plain tensors, not nn.Parameters.
GRAD_MODE = ['all']


class grad_mode:
    """context manager: with grad_mode('arch'): loss = crit(head(x), y); loss.backward()"""

    def __init__(self, mode):
        assert mode in ('all', 'arch', 'weights')
        self.mode = mode

    def __enter__(self):
        self.prev = GRAD_MODE[0]
        GRAD_MODE[0] = self.mode

    def __exit__(self, *a):
        GRAD_MODE[0] = self.prev


def filter_leaves(leaves):
    m = GRAD_MODE[0]
    if m == 'arch':
        return [t for t in leaves if not isinstance(t, torch.nn.Parameter)]
    if m == 'weights':
        return [t for t in leaves if isinstance(t, torch.nn.Parameter)]
    return list(leaves)


class GradArena:
    def __init__(self, tensors, device, flat=None):
        self.tensors = [t for t in tensors]
        self.offsets = {}
        off = 0
        for t in self.tensors:
            self.offsets[id(t)] = (off, t.numel())
            off += (t.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        # flat: a caller-provided buffer of the right size (bmnas.dp puts the arena into peer-mapped symmetric memory)
        assert flat is None or flat.numel() >= max(off, 4)
        self.flat = torch.zeros(max(off, 4), dtype=torch.float32, device=device) if flat is None else flat
        self.views = {id(t): self.flat[o:o + n].view(t.shape) for t, (o, n) in
                      ((t, self.offsets[id(t)]) for t in self.tensors)}
        self.dirty = set()       # ids of leaves whose arena view holds a gradient not yet consumed by step()/zero_grad()
        _ARENAS.add(self)

    def view(self, t):
        return self.views.get(id(t))

    def span(self, tensors):
        """smallest contiguous slice of the flat buffer covering `tensors`"""
        offs = [self.offsets[id(t)] for t in tensors if id(t) in self.offsets]
        if not offs:
            return None
        lo = min(o for o, n in offs)
        hi = max(o + (n + 3) // 4 * 4 for o, n in offs)
        return self.flat[lo:hi]

    def covers(self, tensors):
        return all(id(t) in self.offsets for t in tensors)


def arena_for(root, leaves, device):
    """the root's gradient arena (a parent module may have installed a joint one)"""
    ar = getattr(root, '_bm_arena', None)
    if ar is None or not ar.covers(leaves) or ar.flat.device != device:
        ar = GradArena(leaves, device)
        root._bm_arena = ar
    return ar


class Runner:
    def __init__(self, prog, in_slots, in_need, out, leaves, arena, ptr_sig):
        self.prog = prog
        self.in_slots = in_slots        # slot names of the tensor inputs, in order
        self.in_need = in_need
        self.out = out
        self.leaves = leaves            # tensors whose .grad the backward fills (params + arch)
        self.arena = arena
        self.ptr_sig = ptr_sig
        self.pending_gen = -1

    def forward(self, inputs, masks):
        p = self.prog
        for name, t in zip(self.in_slots, inputs):
            p.bind(name, t)
        if p.use_masks:
            for site, sl in p.mask_slots.items():
                m = masks[site]
                if m.dtype != torch.uint8 or not m.is_contiguous() or m.device != p.device:
                    raise ValueError(f'dropout mask {site}: need a contiguous uint8 tensor on {p.device}')
                p.bind(sl.name, m)
        p.run_forward()
        self.pending_gen = p.generation
        return self.out

    def backward(self, gout, gen):
        p = self.prog
        if gen != p.generation:
            raise RuntimeError('bmnas: stale activation workspace -- another forward of the same module/shape ran '
                               'before this backward (activations live in a static per-shape workspace)')
        p.bind('gout', gout)
        static = STATIC_IO[0]
        # autograd semantics outside STATIC_IO: a leaf that already holds an unconsumed gradient in the arena accumulates
        pending = [] if static else [t for t in self.leaves if t.grad is not None and id(t) in self.arena.dirty
                                     and t.grad.data_ptr() == self.arena.view(t).data_ptr()]
        saved = [t.grad.clone() for t in pending]
        p.run_backward()
        for t, old in zip(pending, saved):
            t.grad.add_(old)
        for t in self.leaves:
            v = self.arena.view(t)
            if t.grad is None:
                t.grad = v
            elif t.grad.data_ptr() != v.data_ptr():
                t.grad.add_(v)
            self.arena.dirty.add(id(t))
        outs = []
        for name, need in zip(self.in_slots, self.in_need):
            g = p._grads.get(name) if need else None
            # inputs the (found) genotype never reads get no gradient
            ok = g is not None and g.data_ptr() in p._written
            outs.append((g.detach() if static else g.detach().clone()) if ok else None)
        return outs


class _ProgFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, masks, n_in, *tensors):
        inputs = tensors[:n_in]
        out = runner.forward(inputs, masks)
        ctx.runner = runner
        ctx.gen = runner.prog.generation
        ctx.n_in = n_in
        ctx.n_rest = len(tensors) - n_in
        ctx.save_for_backward(*inputs)      # keep the bound input pointers alive until backward
        # STATIC_IO: a fresh tensor object over the static output buffer; otherwise a copy the caller may keep
        return out.detach() if STATIC_IO[0] else out.detach().clone()

    @staticmethod
    def backward(ctx, gout):
        gin = ctx.runner.backward(gout.contiguous(), ctx.gen)
        return (None, None, None) + tuple(gin) + (None,) * ctx.n_rest


def _prep(t, device):
    if t.device != device or t.dtype != torch.float32:
        raise ValueError(f'bmnas kernels need fp32 CUDA tensors on {device}; got {t.dtype} on {t.device}')
    return t if t.is_contiguous() else t.contiguous()


def run(root, kind, inputs, build, leaves, C, L, drpt, key_extra=(), masks=None, drop_p=None):
    """Execute (building and caching on first use) the launch plan of `root` for these inputs.

    build(prog, in_slots, need) -> out buffer; must emit the whole forward and register backwards.
    leaves: parameter/architecture tensors that receive gradients.
    """
    if not inputs[0].is_cuda and not N.VALIDATE_ONLY:
        raise N.NativeError('bmnas: the search-step path has no CPU implementation; move the module and its '
                            'inputs to a CUDA device (B200)')
    device = inputs[0].device
    inputs = [_prep(t, device) for t in inputs]
    grad_on = torch.is_grad_enabled()
    need = tuple(bool(t.requires_grad and grad_on) for t in inputs)
    B = inputs[0].shape[0]
    use_masks = masks is not None
    drop_key = tuple(sorted(drop_p.items())) if drop_p else ()
    all_leaves = [t for t in leaves if t.requires_grad]
    leaves = filter_leaves(all_leaves)
    want_backward = bool(grad_on and (any(need) or leaves))
    key = (kind, B, C, L, device.index, bool(root.training), use_masks, need, SAMPLE_OFFSET[0], drop_key,
           GRAD_MODE[0], want_backward) + tuple(key_extra)
    cache = root.__dict__.setdefault('_bm_cache', {})
    ptr_sig = tuple(t.data_ptr() for t in leaves)
    runner = cache.get(key)
    if runner is not None and runner.ptr_sig != ptr_sig:
        runner = None                     # parameters were re-allocated (.to(), new tensors): rebuild
    if runner is None:
        for t in leaves:
            if t.device != device:
                raise ValueError('bmnas: parameter / architecture tensor on the wrong device; call module.to(device)')
        arena = arena_for(root, all_leaves, device)
        prog = Program(device, B, C, L, root.training, drpt)
        prog.use_masks = use_masks
        prog.drop_p = dict(drop_p or {})
        prog.sample_offset = SAMPLE_OFFSET[0]
        prog.want_backward = want_backward
        in_slots = [f'in{i}' for i in range(len(inputs))]
        G = _GradViews(arena, GRAD_MODE[0])
        out = build(prog, [Slot(n) for n in in_slots], need, G)
        prog.seed_grad(out, Slot('gout'))
        if not want_backward:
            prog._stack.clear()            # a no-grad plan keeps nothing for (and emits no) backward
        prog.finalize()
        span = arena.span(leaves)
        if span is not None:
            prog._zero_ranges = [span]
        runner = Runner(prog, in_slots, need, out, leaves, arena, ptr_sig)
        cache[key] = runner
    if any(t.requires_grad for t in inputs) or leaves:
        if grad_on:
            return _ProgFn.apply(runner, masks, len(inputs), *inputs, *leaves)
    out = runner.forward(inputs, masks).detach()
    return out if STATIC_IO[0] else out.clone()


class _GradViews:
    """name -> gradient view lookup used by the emitters: G.get(name) with the P dict alongside"""

    def __init__(self, arena, mode='all'):
        self.arena = arena
        self.P = None
        self.mode = mode

    def _wanted(self, t):
        if t is None or not getattr(t, 'requires_grad', False):
            return False
        if self.mode == 'arch':
            return not isinstance(t, torch.nn.Parameter)
        if self.mode == 'weights':
            return isinstance(t, torch.nn.Parameter)
        return True

    def attach(self, P):
        self.P = P
        return self

    def get(self, name):
        t = self.P.get(name)
        return self.arena.view(t) if self._wanted(t) else None

    def of(self, t):
        return self.arena.view(t) if self._wanted(t) else None


def named_tensors(module, prefix=''):
    """parameters and buffers keyed like state_dict() (live tensors, not copies)"""
    return dict(module.state_dict(prefix=prefix, keep_vars=True))
