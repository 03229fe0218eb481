"""Static launch plans ("programs") for the fusion-cell hot path.

A Program is built once per (shape, mode) and holds
  * every intermediate buffer, pre-allocated in HBM,
  * an ordered list of prepared C-ABI kernel calls for the forward pass and one for
    the backward pass (parameter blocks already filled in),
  * "slots": the few pointers that change per call (user inputs, injected masks).
Running a program is a tight loop of ctypes calls on the current CUDA stream -- no
allocation, no host sync -- so the whole thing is legal inside CUDA-graph capture.

The emitters below restate the reference control flow
  FusionCell.forward   models/search/darts/model_search.py:50-68
  NodeCell.forward     models/search/darts/node_search.py:48-70
  Found_*.forward      models/search/darts/model.py:133-160, node.py:45-76
as kernel sequences; the backward sequence is derived at build time by unwinding
a stack of closures, tracking which gradient buffer has been written so far
(first writer overwrites, later writers accumulate).
"""
import ctypes
import os
import zlib

import torch

from . import native as N

OP_IDS = {'Sum': N.BMNAS_OP_SUM, 'ScaleDotAttn': N.BMNAS_OP_ATTN, 'LinearGLU': N.BMNAS_OP_GLU,
          'ConcatFC': N.BMNAS_OP_FC_RELU, 'CatConvMish': N.BMNAS_OP_FC_MISH}
ATTN_DROP = 0.1          # node_operations.py:89
BN_MOMENTUM, BN_EPS = 0.1, 1e-5


class Slot:
    """A tensor supplied at call time (user input, injected dropout mask, upstream gradient)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f'Slot({self.name})'


SIDE_WGRAD = os.environ.get('BMNAS_SIDE_WGRAD', '1') != '0'   # weight-gradient GEMMs on a side stream (parallel graph branch); see conv_backward
CHAIN_NODE = os.environ.get('BMNAS_CHAIN_NODE', '1') != '0'  # node op i also writes inner edge mix i+1 (CTA-per-sample kernels)
CHAIN_NODE_MAX_B = 640                                       # beyond: the warp-per-sample node kernels, which do not chain
CHAIN_MIX = os.environ.get('BMNAS_CHAIN_MIX', '1') != '0'   # cell-level edge mix + the node cell's first inner mix in one launch
SPLIT_MIX_BWD = os.environ.get('BMNAS_SPLIT_MIX_BWD', '1') != '0'   # edge-mix backward: input grads on the main chain, d(alpha) on the side branch
# NodeMixedOp forward as ONE fused tcgen05 kernel (bmnas_mixed_fwd: conv GEMM + BatchNorm statistics + grid barrier + every
# primitive + gamma-weighted sum, Z written only when a backward will follow) instead of bmnas_conv_fwd + bmnas_node_fwd.
# 0 = off, 1 = whenever the shape is supported, default: batches of FUSED_MIXED_MIN_B samples or more
FUSED_MIXED = os.environ.get('BMNAS_FUSED_MIXED', 'auto')
FUSED_MIXED_MIN_B = int(os.environ.get('BMNAS_FUSED_MIXED_MIN_B', '768'))   # measured crossover (profiles/r02_fused_crossover.txt): 3xTF32 ties at 512, wins from 1024
# ... and below FUSED_MIXED_MIN_B the small-batch variant (bmnas_mixed_small_fwd: fp32 FFMA tiles, one grid barrier, also writes the
# chained inner edge mix): 0 = off (bmnas_conv_fwd + bmnas_node_fwd), 1 = whenever the library takes the shape
FUSED_MIXED_SMALL = os.environ.get('BMNAS_FUSED_MIXED_SMALL', '1')
FUSED_OPS_RANK = {'Sum': 0, 'ScaleDotAttn': 1, 'LinearGLU': 2, 'ConcatFC': 3, 'CatConvMish': 3}
_side_streams = {}
# SearchStep's half steps: the gradient-arena span a plan accumulates into is cleared at the START of the forward, on the
# side branch, instead of at the start of the backward on the main chain (where the memset node and its edges cost ~10 us
# between the loss and the first backward kernel, tools/timeline.py).  Legal there because the optimiser consumed the span
# at the end of the previous half step and forward + backward are always issued together; plain autograd use of the
# modules keeps the zero fill in the backward (pending gradients may live in the span until then).
EARLY_ZERO = [False]
EARLY_ZERO_AT = int(os.environ.get('BMNAS_EARLY_ZERO_AT', '4'))   # forward launch after which the clear is forked


class early_zero:
    def __enter__(self):
        self.prev = EARLY_ZERO[0]
        EARLY_ZERO[0] = True

    def __exit__(self, *a):
        EARLY_ZERO[0] = self.prev



def side_stream(device):
    st = _side_streams.get(device)
    if st is None:
        st = _side_streams[device] = torch.cuda.Stream(device=device)
    return st


def join_side(device):
    """the current stream waits for everything forked onto the side branch (no-op if it was never used)"""
    st = _side_streams.get(device)
    if st is None:
        return
    cur = torch.cuda.current_stream(device)
    if torch.cuda.is_current_stream_capturing():
        with torch.cuda.stream(st):
            forked = torch.cuda.is_current_stream_capturing()
        if not forked:        # nothing of this capture runs on the side branch: an edge to it would be illegal
            return
    cur.wait_stream(st)


def uid_of(name):
    return zlib.crc32(name.encode()) & 0xffffffff


class Program:
    def __init__(self, device, B, C, L, training, drpt):
        self.device = device
        self.B, self.C, self.L = B, C, L
        self.training = bool(training)
        self.drpt = float(drpt)
        self.fwd, self.bwd = [], []
        self._cur = self.fwd
        self._stack = []
        self._bindings = {}
        self._keep = []
        self._grads = {}         # id(state tensor) or slot name -> grad buffer
        self._written = set()    # data_ptrs of grad buffers already written in the backward order
        self._zero_ranges = []   # (tensor) zeroed at the start of every backward
        self.drop_p = {}         # dropout site name -> p (read from the nn.Dropout modules)
        self.mask_slots = {}     # dropout site name -> Slot (parity mode)
        self.use_masks = False
        self.rng_state = None
        self.sample_offset = 0
        self.outputs = {}
        self.generation = 0
        self._prep = []          # convs whose weight images bmnas_wprep refreshes at the start of every forward
        self._prep_calls = []
        self._rng_in_prep = False
        self.n_fwd_launches = 0
        self.n_bwd_launches = 0
        self.want_backward = True    # False: a no-grad forward (metrics pass, inference): nothing is kept for a backward
        self._cur_tag = None
        # one-graph search step (bmnas.search): the weight half's bmnas_wprep runs on the side branch of the ARCH half (the
        # Architect does not touch the weights), so it leaves the critical path of the step
        self.side_extra = None       # callable(stream_ptr): extra work forked next to the gradient-span clear
        self._prep_hoisted = None    # event: this plan's weight images (and Philox step) were already advanced elsewhere

    def fused_mixed_ok(self, ops, alias=True):
        """shape-level test for the fused NodeMixedOp forward kernel (the library re-checks pointers / alignment)"""
        if FUSED_MIXED == '0' or not alias or self.C != 128 or self.L not in (4, 8, 16):
            return False
        if FUSED_MIXED == 'auto' and self.B < FUSED_MIXED_MIN_B:
            return False
        if N.lib().bmnas_get_gemm_mode() == 0:
            return False
        ranks = [FUSED_OPS_RANK.get(o, -1) for o in ops]
        return (all(r >= 0 for r in ranks) and ranks == sorted(set(ranks)) and 'LinearGLU' in ops
                and ops[-1] in ('ConcatFC', 'CatConvMish') and ops.index('LinearGLU') == len(ops) - 2)

    def fused_small_ok(self, ops):
        """shape-level test for bmnas_mixed_small_fwd (csrc/mixed_small.cu)"""
        C, L, B = self.C, self.L, self.B
        if C % 32 or C > 256 or L not in (4, 8, 16):
            return False
        if ((B * L + 31) // 32) * (C // 32) > 148:          # one co-resident wave (grid barrier)
            return False
        if int(N.lib().bmnas_conv_image_fmt(B, L, C, 3 * C)) != 1:   # the library would put this GEMM on the tensor cores
            return False
        ranks = [FUSED_OPS_RANK.get(o, -1) for o in ops]
        return (all(r >= 0 for r in ranks) and ranks == sorted(set(ranks)) and 'LinearGLU' in ops
                and ops[-1] in ('ConcatFC', 'CatConvMish') and ops.index('LinearGLU') == len(ops) - 2)

    # ------------------------------------------------------------------ storage
    def buf(self, *shape, dtype=torch.float32, zero=False):
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
        self._keep.append(t)
        return t

    def counter(self, n=1):
        return self.buf(max(int(n), 1), dtype=torch.int32, zero=True)

    def ensure_rng(self):
        if self.rng_state is None:
            from . import rng
            self.rng_state = self.buf(2, dtype=torch.int64, zero=True)
            self.rng_state[0] = rng.next_seed()
        return self.rng_state

    def grad_of(self, t):
        key = t.name if isinstance(t, Slot) else id(t)
        g = self._grads.get(key)
        if g is None:
            g = self.buf(*t.shape) if torch.is_tensor(t) else self.buf(self.B, self.C, self.L)
            self._grads[key] = g
        return g

    def out_grad(self, slot, buf):
        """`buf` receives the gradient of the call-time input `slot` (written by a kernel directly)"""
        self._grads[slot.name] = buf
        self._written.add(buf.data_ptr())

    def seed_grad(self, t, slot):
        """the gradient of program output t arrives from outside through `slot`"""
        self._grads[id(t)] = slot

    def has_grad(self, t):
        key = t.name if isinstance(t, Slot) else id(t)
        g = self._grads.get(key)
        if isinstance(g, Slot):
            return True
        return g is not None and g.data_ptr() in self._written

    def acc(self, g):
        """accumulate flag for grad buffer g in the backward order being emitted"""
        if g is None:
            return 0
        assert not isinstance(g, Slot), 'cannot accumulate into an upstream-gradient slot'
        k = g.data_ptr()
        if k in self._written:
            return 1
        self._written.add(k)
        return 0

    # ------------------------------------------------------------------ struct plumbing
    def setp(self, st, field, val, idx=None, offset=0):
        if isinstance(val, Slot):
            self._bindings.setdefault(val.name, []).append((st, field, idx, offset))
            p = None
        elif val is None:
            p = None
        else:
            self._keep.append(val)
            p = val.data_ptr() + offset
        if idx is None:
            setattr(st, field, p)
        else:
            getattr(st, field)[idx] = p

    def bind(self, name, tensor):
        base = tensor.data_ptr()
        for st, field, idx, off in self._bindings.get(name, ()):
            if idx is None:
                setattr(st, field, base + off)
            else:
                getattr(st, field)[idx] = base + off

    def emit(self, name, st, side=False, args=None):
        self._cur.append(N.Call(name, st, side=side, args=args, tag=self._cur_tag))

    def on_backward(self, fn):
        self._stack.append(fn)

    def finalize(self):
        """derive the backward launch list by unwinding the closure stack"""
        self._cur = self.bwd
        while self._stack:
            self._stack.pop()()
        self._cur = None
        # programmatic dependent launch: a conv may fetch its weight image before griddepcontrol.wait when at least
        # one kernel sits between bmnas_wprep and it (include/bmnas_b200.h, early_ok)
        for i, c in enumerate(self.fwd):
            if c.name == 'bmnas_conv_fwd' and i >= 1:
                c.st.early_ok = 1
        for c in self.bwd:
            if c.name == 'bmnas_conv_dgrad':
                c.st.early_ok = 1
        for i0 in range(0, len(self._prep), N.BMNAS_MAX_PREP):
            group = self._prep[i0:i0 + N.BMNAS_MAX_PREP]
            st = N.bmnas_wprep_params()
            st.n = len(group)
            q = 0
            for i, c in enumerate(group):
                st.M[i], st.K[i], st.w_fold[i], st.n_seg[i] = c['M'], c['K'], c['fold'], len(c['segs'])
                for j, (W, m) in enumerate(c['segs']):
                    st.seg_M[i * N.BMNAS_MAX_SEG + j] = m
                    self.setp(st, 'W', W, i * N.BMNAS_MAX_SEG + j)
                self.setp(st, 'img_fwd', c['img_f'], i)
                self.setp(st, 'img_dgrad', c['img_d'], i)
                st.fmt[i] = c['fmt']
                st.q_start[i] = q
                q += int(N.lib().bmnas_wprep_items(c['M'], c['K'], c['fmt']))
            st.q_start[len(group)] = q
            self._prep_calls.append(N.Call('bmnas_wprep', st))
        # the first weight-image launch also advances the dropout step counter (one launch less per forward)
        self._rng_in_prep = bool(self._prep_calls) and self.rng_state is not None
        if self._rng_in_prep:
            self.setp(self._prep_calls[0].st, 'rng_state', self.rng_state)
        self.n_fwd_launches = len(self.fwd) + len(self._prep_calls) + (1 if (self.rng_state is not None and not self._rng_in_prep) else 0)
        self.n_bwd_launches = len(self.bwd) + len(self._zero_ranges)

    # ------------------------------------------------------------------ execution
    def _zero(self, stream_ptr):
        for t in self._zero_ranges:
            N.launch('bmnas_zero', ctypes.c_void_p(t.data_ptr()), ctypes.c_longlong(t.numel() * t.element_size()), stream_ptr)

    def run_prep(self, s):
        """the per-forward preamble: Philox step counter + weight images (bmnas_wprep) on stream pointer s"""
        if self.rng_state is not None and not self._rng_in_prep:
            N.launch('bmnas_rng_advance', ctypes.c_void_p(self.rng_state.data_ptr()), s)
        for c in self._prep_calls:
            c(s)

    def run_forward(self):
        s = N.current_stream()
        self._zero_ev = None
        early = EARLY_ZERO[0] and self._zero_ranges and self.want_backward and SIDE_WGRAD and not N.VALIDATE_ONLY
        fork_at = min(EARLY_ZERO_AT, len(self.fwd) - 1) if early else -1
        if self._prep_hoisted is not None:
            torch.cuda.current_stream().wait_event(self._prep_hoisted)
            self._prep_hoisted = None
        else:
            self.run_prep(s)
        for i, c in enumerate(self.fwd):
            c(s)
            if i == fork_at:
                # not at the very start: a fork right behind the root of a captured graph delayed the first kernels of
                # the main chain by 5-11 us (tools/timeline.py)
                main = torch.cuda.current_stream()
                side = side_stream(self.device)
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                self._zero(ctypes.c_void_p(side.cuda_stream))
                self._zero_ev = torch.cuda.Event()
                self._zero_ev.record(side)
                if self.side_extra is not None:
                    self.side_extra(ctypes.c_void_p(side.cuda_stream), side)
        self.generation += 1

    def run_backward(self):
        s = N.current_stream()
        if getattr(self, '_zero_ev', None) is not None:
            torch.cuda.current_stream().wait_event(self._zero_ev)      # cleared during the forward, on the side branch
            self._zero_ev = None
        else:
            self._zero(s)
        if not SIDE_WGRAD or N.VALIDATE_ONLY:
            for c in self.bwd:
                c(s)
            return
        main = torch.cuda.current_stream()
        side = side_stream(self.device)
        sp = ctypes.c_void_p(side.cuda_stream)
        forked = False
        for c in self.bwd:
            if c.side:
                ev = c.keep
                if not ev:
                    ev = c.keep = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                c(sp)
                forked = True
            else:
                c(s)
        if forked:
            main.wait_stream(side)

    # ------------------------------------------------------------------ dropout helper
    def p_of(self, site, default):
        return float(self.drop_p.get(site, default))

    def _drop(self, site, p):
        """(mask Slot or None, needs rng) for a dropout site"""
        if not self.training or p <= 0.0:
            return None
        if self.use_masks:
            sl = self.mask_slots.get(site)
            if sl is None:
                sl = Slot('mask:' + site)
                self.mask_slots[site] = sl
            return sl
        self.ensure_rng()
        return None

    # ------------------------------------------------------------------ kernels: edge mix
    def mix(self, xs, w, w_off, logits, out, gw=None, need=None, chain=None, dots_only=False):
        """out = sum_j w[w_off+j, skip] * xs[j]; registers the backward.
        chain = dict(w=, w_off=, n=, out=): a second output chain['out'] = (sum of the n skip weights of
        chain['w'] from row w_off) * out, written by the same launch -- the first inner edge mix of a searchable
        NodeCell, which reads this mix's output twice; the backward then takes gout + s2 * g(chain out).
        dots_only: that inner mix itself -- no forward launch, no input-gradient launch (both folded into the
        producer's), only its d(beta) dot products on the side branch."""
        xs = list(xs)
        n = len(xs)
        assert 1 <= n <= N.BMNAS_MAX_MIX
        numel = self.B * self.C * self.L

        def set_chain(st, field, t):
            st.n2 = chain['n']
            self.setp(st, 'w2', chain['w'], offset=chain['w_off'] * 8)
            self.setp(st, field, t)
        if not dots_only:
            st = N.bmnas_mix_params()
            st.n, st.w_is_logits, st.numel = n, int(logits), numel
            for j, x in enumerate(xs):
                self.setp(st, 'x', x, j)
            self.setp(st, 'w', w, offset=w_off * 8)
            self.setp(st, 'out', out)
            if chain is not None:
                set_chain(st, 'out2', chain['out'])
            self.emit('bmnas_mix_fwd', st)
        need = need if need is not None else [True] * n

        def bwd():
            g2 = chain['out'] if (chain is not None and self.has_grad(chain['out'])) else None
            if not self.has_grad(out) and g2 is None:
                return
            gout = self.grad_of(out) if self.has_grad(out) else None

            def base():
                sb = N.bmnas_mix_params()
                sb.n, sb.w_is_logits, sb.numel = n, int(logits), numel
                for j, x in enumerate(xs):
                    self.setp(sb, 'x', x, j)
                self.setp(sb, 'w', w, offset=w_off * 8)
                self.setp(sb, 'gout', gout)
                if g2 is not None:
                    set_chain(sb, 'gout2', self.grad_of(g2))
                return sb
            # the input gradients are what the rest of the backward waits for: a pure streaming launch on the main
            # chain.  d(alpha) (dot products + cross-CTA reduction) only feeds the optimiser, so it goes on the side
            # branch next to the weight-gradient GEMMs.
            sb = base()
            any_gx = False
            if not dots_only:
                for j, x in enumerate(xs):
                    if need[j]:
                        g = self.grad_of(x)
                        sb.gx_accum[j] = self.acc(g)
                        self.setp(sb, 'gx', g, j)
                        any_gx = True
            split = (SPLIT_MIX_BWD or dots_only) and gw is not None and (any_gx or dots_only)
            if gw is not None and not split:
                self.setp(sb, 'gw', gw, offset=w_off * 8)
                self.setp(sb, 'partials', self.buf(int(N.lib().bmnas_mix_partials_size(ctypes.byref(sb))), zero=True))
                self.setp(sb, 'counter', self.counter())
            if any_gx or (gw is not None and not split):
                self.emit('bmnas_mix_bwd', sb)
            if split:
                sd = base()
                self.setp(sd, 'gw', gw, offset=w_off * 8)
                self.setp(sd, 'partials', self.buf(int(N.lib().bmnas_mix_partials_size(ctypes.byref(sd))), zero=True))
                self.setp(sd, 'counter', self.counter())
                self.emit('bmnas_mix_bwd', sd, side=True)
        self.on_backward(bwd)

    # ------------------------------------------------------------------ kernels: conv (+BN stats) and its backward
    def conv(self, srcs, src_C, segs, w_fold, bn, emit=True, fwd_fmt=None, want_Z=True):
        """segs: list of dict(W=, bias=, M=, rm=, rv=, nbt=, gW=, gbias=).  Returns dict(Z, mean, rstd, M, st).
        emit=False: the caller launches the forward itself (fused mixed op); fwd_fmt: image format of the forward weight
        image (the dgrad image keeps the library's choice); want_Z=False: no Z buffer (no backward will follow)."""
        st = N.bmnas_conv_params()
        K = sum(src_C)
        M = sum(s['M'] for s in segs)
        st.B, st.L, st.K, st.w_fold, st.n_src, st.n_seg, st.M = self.B, self.L, K, w_fold, len(srcs), len(segs), M
        st.bn_mode = (1 if self.training else 2) if bn else 0
        st.momentum, st.eps = BN_MOMENTUM, BN_EPS
        for i, (s, c) in enumerate(zip(srcs, src_C)):
            self.setp(st, 'src', s, i)
            st.src_C[i] = c
        for i, sg in enumerate(segs):
            st.seg_M[i] = sg['M']
            self.setp(st, 'W', sg['W'], i)
            self.setp(st, 'bias', sg.get('bias'), i)
            if bn:
                self.setp(st, 'running_mean', sg['rm'], i)
                self.setp(st, 'running_var', sg['rv'], i)
                self.setp(st, 'num_batches_tracked', sg['nbt'], i)
        Z = self.buf(self.B, M, self.L) if want_Z else None
        self.setp(st, 'Z', Z)
        # weight images (refreshed by ONE bmnas_wprep launch at the start of every forward): the library picks the
        # format = GEMM engine for this problem size (plain fp32 for the small-N cp.async kernels, tcgen05 slabs beyond)
        img_f = img_d = None
        fmt = int(N.lib().bmnas_conv_image_fmt_dgrad(self.B, self.L, K, M))        # engine of the dgrad GEMM
        fmt_f = int(N.lib().bmnas_conv_image_fmt(self.B, self.L, K, M)) if fwd_fmt is None else fwd_fmt
        if fmt_f >= 0 and all(c % 4 == 0 for c in src_C) and all(sg['W'].data_ptr() % 16 == 0 for sg in segs):
            seg_list = [(sg['W'], sg['M']) for sg in segs]
            img_f = self.buf(int(N.lib().bmnas_wimg_floats_fmt(M, K, 0, fmt_f)))
            self.setp(st, 'wimg_fwd', img_f)
            st.wimg_fmt = fmt_f
            if fmt >= 0 and want_Z:
                img_d = self.buf(int(N.lib().bmnas_wimg_floats_fmt(M, K, 1, fmt)))
            if fmt == fmt_f:
                self._prep.append(dict(M=M, K=K, fold=w_fold, segs=seg_list, img_f=img_f, img_d=img_d, fmt=fmt))
            else:       # two formats: the conv is listed twice, each entry writes one image (NULL = skipped)
                self._prep.append(dict(M=M, K=K, fold=w_fold, segs=seg_list, img_f=img_f, img_d=None, fmt=fmt_f))
                if img_d is not None:
                    self._prep.append(dict(M=M, K=K, fold=w_fold, segs=seg_list, img_f=None, img_d=img_d, fmt=fmt))
        mean = rstd = None
        if bn:
            mean, rstd = self.buf(M), self.buf(M)
            self.setp(st, 'mean', mean)
            self.setp(st, 'rstd', rstd)
            if self.training:
                self.setp(st, 'stat_part', self.buf(int(N.lib().bmnas_conv_stat_part_size(ctypes.byref(st)))))
                self.setp(st, 'counter', self.counter(N.lib().bmnas_conv_num_counters(ctypes.byref(st))))
        if emit:
            self.emit('bmnas_conv_fwd', st)
        return dict(Z=Z, mean=mean, rstd=rstd, M=M, K=K, srcs=srcs, src_C=src_C, segs=segs, w_fold=w_fold, img_d=img_d,
                    fmt=max(fmt, 0), st=st)

    def conv_backward(self, cv, GV, coef, need_src):
        """dgrad into the source grads + wgrad into the parameter grad views."""
        srcs, src_C, segs = cv['srcs'], cv['src_C'], cv['segs']

        def base():
            st = N.bmnas_conv_params()
            st.B, st.L, st.K, st.w_fold, st.n_src, st.n_seg, st.M = (self.B, self.L, cv['K'], cv['w_fold'], len(srcs),
                                                                    len(segs), cv['M'])
            for i, (s, c) in enumerate(zip(srcs, src_C)):
                self.setp(st, 'src', s, i)
                st.src_C[i] = c
            for i, sg in enumerate(segs):
                st.seg_M[i] = sg['M']
                self.setp(st, 'W', sg['W'], i)
            self.setp(st, 'GV', GV)
            self.setp(st, 'Z', cv['Z'])
            self.setp(st, 'wimg_dgrad', cv.get('img_d'))
            st.wimg_fmt = cv.get('fmt', 0)
            if coef is not None:
                self.setp(st, 'coef_a', coef[0])
                self.setp(st, 'coef_b', coef[1])
                self.setp(st, 'coef_c', coef[2])
            return st

        # wgrad first: nothing downstream reads the parameter gradients before the optimiser, so
        # run_backward() forks it onto a side stream right behind the kernel that produced GV/coef and the
        # main chain continues with dgrad (a parallel branch of the captured CUDA graph)
        if any(sg.get('gW') is not None or sg.get('gbias') is not None for sg in segs):
            st = base()
            for i, sg in enumerate(segs):
                self.setp(st, 'gW', sg.get('gW'), i)
                self.setp(st, 'gbias', sg.get('gbias'), i)
            self.emit('bmnas_conv_wgrad', st, side=True)
        if any(need_src):
            st = base()
            for i, s in enumerate(srcs):
                if need_src[i]:
                    g = self.grad_of(s)
                    st.gsrc_accum[i] = self.acc(g)
                    self.setp(st, 'gsrc', g, i)
            self.emit('bmnas_conv_dgrad', st)

    # ------------------------------------------------------------------ kernels: step-node mixed op
    def node_op(self, x, y, ops, P, G, prefix_of, gamma, gamma_off, logits, out, g_gamma=None,
                need_x=True, need_y=True, conv_srcs=None, conv_src_C=None, conv_need=None, chain=None):
        """ops: list of primitive names; prefix_of(k) -> parameter prefix of op k.
        x, y: tensors/Slots (x is y => aliased).  Emits conv (if any conv-backed op) + node kernels.
        conv_srcs / conv_src_C / conv_need: the conv reads these sources (any channel counts) instead of cat(x, y)
        -- the reshape layers, whose conv block is ConcatFC over one pooled source; x and y are then unused.
        chain = dict(w=, w_off=, priors=[...], need=[...], out=): the NEXT inner edge mix, written by this op's
        forward launch (out2) and folded into its backward (see bmnas_node_params in the header)."""
        C, L = self.C, self.L
        alias = x is y
        ext = conv_srcs is not None
        if ext:
            assert all(n in ('ConcatFC', 'CatConvMish', 'LinearGLU') for n in ops)
            need_x = need_y = False
        segs, z_off, off = [], {}, 0
        for k, name in enumerate(ops):
            if name in ('LinearGLU', 'ConcatFC', 'CatConvMish'):
                rows = 2 * C if name == 'LinearGLU' else C
                pre = prefix_of(k)
                segs.append(dict(W=P[pre + '.conv.weight'], bias=P[pre + '.conv.bias'], M=rows,
                                 rm=P[pre + '.bn.running_mean'], rv=P[pre + '.bn.running_var'],
                                 nbt=P[pre + '.bn.num_batches_tracked'],
                                 gW=G.get(pre + '.conv.weight'), gbias=G.get(pre + '.conv.bias')))
                z_off[k] = off
                off += rows
        assert len(segs) <= N.BMNAS_MAX_SEG
        cv = None
        self._mixed_id = getattr(self, '_mixed_id', 0) + 1
        prev_tag, self._cur_tag = self._cur_tag, f'mixed{self._mixed_id}'
        fused = bool(segs) and not ext and chain is None and self.fused_mixed_ok(ops, alias)
        # small batches: the FFMA variant (takes the chained mix too); decided on the shape, the library re-checks pointers
        small = (not fused and bool(segs) and not ext and alias and FUSED_MIXED_SMALL != '0' and self.fused_small_ok(ops))
        if segs and ext:
            cv = self.conv(list(conv_srcs), list(conv_src_C), segs, 1, bn=True)
            x = y = cv['Z']               # placeholders: no primitive of an external-source op reads x / y
            alias = True
        elif segs:
            cv = self.conv([x] if alias else [x, y], [C] if alias else [C, C], segs, 2 if alias else 1, bn=True,
                           emit=not (fused or small),
                           fwd_fmt=((2 if N.lib().bmnas_get_gemm_mode() == 3 else 0) if fused else (1 if small else None)),
                           want_Z=(self.want_backward or not (fused or small)))
        M = cv['M'] if cv else 0

        def fill(st):
            st.B, st.C, st.L, st.n_ops, st.M = self.B, C, L, len(ops), M
            st.training, st.gamma_is_logits, st.alias_xy = int(self.training), int(logits), int(alias)
            st.sample_offset = self.sample_offset
            self.setp(st, 'x', x)
            self.setp(st, 'y', y)
            if cv:
                self.setp(st, 'Z', cv['Z'])
                self.setp(st, 'mean', cv['mean'])
                self.setp(st, 'rstd', cv['rstd'])
            if gamma is not None:
                self.setp(st, 'gamma', gamma, offset=gamma_off * 4)
            for k, name in enumerate(ops):
                pre = prefix_of(k)
                st.op_type[k] = OP_IDS[name]
                st.z_off[k] = z_off.get(k, 0)
                p = 0.0 if name == 'Sum' else self.p_of(pre + '.dropout', ATTN_DROP if name == 'ScaleDotAttn' else self.drpt)
                st.p_drop[k] = p
                st.op_uid[k] = uid_of(pre)
                if name != 'Sum':
                    self.setp(st, 'mask', self._drop(pre + '.dropout', p), k)
                if name == 'ScaleDotAttn':
                    self.setp(st, 'ln_w', P[pre + '.ln.weight'], k)
                    self.setp(st, 'ln_b', P[pre + '.ln.bias'], k)
                elif name != 'Sum':
                    self.setp(st, 'bn_w', P[pre + '.bn.weight'], k)
                    self.setp(st, 'bn_b', P[pre + '.bn.bias'], k)
            if self.rng_state is not None:
                self.setp(st, 'rng_state', self.rng_state)

        def fill_chain(st):
            st.n_chain = len(chain['priors'])
            st.chain_is_logits = int(logits)
            self.setp(st, 'chain_w', chain['w'], offset=chain['w_off'] * 8)
            for j, t in enumerate(chain['priors']):
                self.setp(st, 'chain_x', t, j)

        st = N.bmnas_node_params()
        fill(st)
        self.setp(st, 'out', out)
        if chain is not None:
            assert len(chain['priors']) <= N.BMNAS_MAX_SRC
            fill_chain(st)
            self.setp(st, 'out2', chain['out'])
        st.early_ok = 1 if cv else 0      # the conv GEMM sits between the producers of x / y and this kernel
        if fused and N.lib().bmnas_mixed_supported(ctypes.byref(cv['st']), ctypes.byref(st)):
            ws = self.buf((int(N.lib().bmnas_mixed_workspace_bytes()) + 3) // 4, zero=True)
            self.emit('bmnas_mixed_fwd', st, args=(ctypes.byref(cv['st']), ctypes.byref(st), ctypes.c_void_p(ws.data_ptr())))
            self._keep.append(cv['st'])
        elif small and N.lib().bmnas_mixed_small_supported(ctypes.byref(cv['st']), ctypes.byref(st)):
            nbytes = int(N.lib().bmnas_mixed_small_workspace_bytes(ctypes.byref(cv['st']), ctypes.byref(st)))
            ws = self.buf((nbytes + 3) // 4, zero=True)
            # the weight tile may be fetched before griddepcontrol.wait when a kernel sits between bmnas_wprep and this one
            cv['st'].early_ok = 1 if len(self.fwd) >= 1 else 0
            self.emit('bmnas_mixed_small_fwd', st, args=(ctypes.byref(cv['st']), ctypes.byref(st), ctypes.c_void_p(ws.data_ptr())))
            self._keep.append(cv['st'])
        else:
            if fused or small:            # the library declined (alignment / pointers): the two-kernel path
                assert cv['Z'] is not None, 'fused mixed op declined by the library in a no-grad plan'
                self.emit('bmnas_conv_fwd', cv['st'])
            self.emit('bmnas_node_fwd', st)
        self._cur_tag = prev_tag

        tag = f'mixed{self._mixed_id}'

        def bwd():
            g2 = chain['out'] if (chain is not None and self.has_grad(chain['out'])) else None
            if not self.has_grad(out) and g2 is None:
                return
            self._cur_tag = tag
            sb = N.bmnas_node_params()
            fill(sb)
            sb.early_ok = 1                   # backward: x, y, Z, mean, rstd are forward tensors
            self.setp(sb, 'gout', self.grad_of(out) if self.has_grad(out) else None)
            if g2 is not None:
                fill_chain(sb)
                self.setp(sb, 'gout2', self.grad_of(g2))
                for j, t in enumerate(chain['priors']):
                    if chain['need'][j]:
                        g = self.grad_of(t)
                        sb.chain_gx_accum[j] = self.acc(g)
                        self.setp(sb, 'chain_gx', g, j)
            if need_x:
                gx = self.grad_of(x)
                sb.gx_accum = self.acc(gx)
                self.setp(sb, 'gx', gx)
            if need_y and not alias:
                gy = self.grad_of(y)
                sb.gy_accum = self.acc(gy)
                self.setp(sb, 'gy', gy)
            coef = GV = None
            if cv:
                GV = self.buf(self.B, M, L)
                coef = (self.buf(M), self.buf(M), self.buf(M))
                self.setp(sb, 'GV', GV)
                self.setp(sb, 'coef_a', coef[0])
                self.setp(sb, 'coef_b', coef[1])
                self.setp(sb, 'coef_c', coef[2])
            if g_gamma is not None:
                self.setp(sb, 'g_gamma', g_gamma, offset=gamma_off * 4)
            for k, name in enumerate(ops):
                pre = prefix_of(k)
                if name == 'ScaleDotAttn':
                    self.setp(sb, 'g_ln_w', G.get(pre + '.ln.weight'), k)
                    self.setp(sb, 'g_ln_b', G.get(pre + '.ln.bias'), k)
                elif name != 'Sum':
                    self.setp(sb, 'g_bn_w', G.get(pre + '.bn.weight'), k)
                    self.setp(sb, 'g_bn_b', G.get(pre + '.bn.bias'), k)
            self.setp(sb, 'partials', self.buf(int(N.lib().bmnas_node_partials_size(ctypes.byref(sb))), zero=True))
            self.setp(sb, 'counter', self.counter())
            self.emit('bmnas_node_bwd', sb)
            if cv:
                self.conv_backward(cv, GV, coef, list(conv_need) if ext else ([need_x] if alias else [need_x, need_y]))
            self._cur_tag = None
        self.on_backward(bwd)

    # ------------------------------------------------------------------ kernels: adaptive max pool (reshape layers)
    def pool(self, x, Cin, H, W, OH, OW, need_x):
        """pooled (B, Cin, OH*OW) = AdaptiveMaxPool2d((OH, OW)) of the raw feature x (B, Cin, H, W)
        (aux_models.py:61-69, 102-110); registers the gather backward when the input wants a gradient."""
        pooled = self.buf(self.B, Cin, OH * OW)
        st = N.bmnas_pool_params()
        st.B, st.C, st.H, st.W, st.OH, st.OW = self.B, Cin, H, W, OH, OW
        self.setp(st, 'x', x)
        self.setp(st, 'out', pooled)
        idx = None
        if need_x:
            idx = self.buf(self.B, Cin, OH * OW, dtype=torch.int32)
            self.setp(st, 'argmax', idx)
        self.emit('bmnas_pool_fwd', st)

        def bwd():
            if not need_x or not self.has_grad(pooled):
                return
            sb = N.bmnas_pool_params()
            sb.B, sb.C, sb.H, sb.W, sb.OH, sb.OW = self.B, Cin, H, W, OH, OW
            gx = self.buf(self.B, Cin, H, W)
            self._grads[x.name if isinstance(x, Slot) else id(x)] = gx
            sb.gx_accum = self.acc(gx)
            self.setp(sb, 'gout', self._grads[id(pooled)])
            self.setp(sb, 'argmax', idx)
            self.setp(sb, 'gx', gx)
            self.emit('bmnas_pool_bwd', sb)
        self.on_backward(bwd)
        return pooled

    # ------------------------------------------------------------------ kernels: LayerNorm blocks
    def ln_cat(self, srcs, src_C, residual, ln_w, ln_b, g_ln_w, g_ln_b, relu, out, need_src=None, need_res=True):
        Ctot = sum(src_C)

        def fill(st):
            st.B, st.L, st.Ctot, st.n_src, st.mode, st.relu_out = self.B, self.L, Ctot, len(srcs), 0, int(relu)
            st.training = int(self.training)
            for i, (s, c) in enumerate(zip(srcs, src_C)):
                self.setp(st, 'src', s, i)
                st.src_C[i] = c
            self.setp(st, 'residual', residual)
            self.setp(st, 'ln_w', ln_w)
            self.setp(st, 'ln_b', ln_b)
        st = N.bmnas_ln_params()
        fill(st)
        self.setp(st, 'out', out)
        self.emit('bmnas_ln_fwd', st)
        need_src = need_src if need_src is not None else [True] * len(srcs)

        def bwd():
            if not self.has_grad(out):
                return
            sb = N.bmnas_ln_params()
            fill(sb)
            self.setp(sb, 'gout', self.grad_of(out))
            for i, s in enumerate(srcs):
                if need_src[i]:
                    g = self.grad_of(s)
                    sb.gsrc_accum[i] = self.acc(g)
                    self.setp(sb, 'gsrc', g, i)
            if residual is not None and need_res:
                g = self.grad_of(residual)
                sb.gres_accum = self.acc(g)
                self.setp(sb, 'gresidual', g)
            self.setp(sb, 'g_ln_w', g_ln_w)
            self.setp(sb, 'g_ln_b', g_ln_b)
            self.emit('bmnas_ln_bwd', sb)
        self.on_backward(bwd)

    def ln_tail(self, cv, residual, bn_w, bn_b, g_bn_w, g_bn_b, ln_w, ln_b, g_ln_w, g_ln_b, site, out,
                need_src, need_res=True):
        C = self.C
        p_tail = self.p_of(site, self.drpt)
        mask = self._drop(site, p_tail)

        def fill(st):
            st.B, st.L, st.Ctot, st.n_src, st.mode, st.relu_out = self.B, self.L, C, 1, 1, 0
            st.training, st.p_drop, st.op_uid = int(self.training), p_tail, uid_of(site)
            st.sample_offset = self.sample_offset
            st.src_C[0] = C
            self.setp(st, 'src', cv['Z'], 0)
            self.setp(st, 'residual', residual)
            self.setp(st, 'mean', cv['mean'])
            self.setp(st, 'rstd', cv['rstd'])
            self.setp(st, 'bn_w', bn_w)
            self.setp(st, 'bn_b', bn_b)
            self.setp(st, 'mask', mask)
            if self.rng_state is not None:
                self.setp(st, 'rng_state', self.rng_state)
            self.setp(st, 'ln_w', ln_w)
            self.setp(st, 'ln_b', ln_b)
        st = N.bmnas_ln_params()
        fill(st)
        self.setp(st, 'out', out)
        self.emit('bmnas_ln_fwd', st)

        def bwd():
            if not self.has_grad(out):
                return
            sb = N.bmnas_ln_params()
            fill(sb)
            self.setp(sb, 'gout', self.grad_of(out))
            GV = self.buf(self.B, C, self.L)
            coef = (self.buf(C), self.buf(C), self.buf(C))
            self.setp(sb, 'gsrc', GV, 0)
            if need_res:
                g = self.grad_of(residual)
                sb.gres_accum = self.acc(g)
                self.setp(sb, 'gresidual', g)
            self.setp(sb, 'g_ln_w', g_ln_w)
            self.setp(sb, 'g_ln_b', g_ln_b)
            self.setp(sb, 'g_bn_w', g_bn_w)
            self.setp(sb, 'g_bn_b', g_bn_b)
            self.setp(sb, 'coef_a', coef[0])
            self.setp(sb, 'coef_b', coef[1])
            self.setp(sb, 'coef_c', coef[2])
            self.setp(sb, 'partials', self.buf(int(N.lib().bmnas_ln_partials_size(ctypes.byref(sb))), zero=True))
            self.setp(sb, 'counter', self.counter())
            self.emit('bmnas_ln_bwd', sb)
            self.conv_backward(cv, GV, coef, need_src)
        self.on_backward(bwd)

    # ------------------------------------------------------------------ composite: node cell tail
    def node_tail(self, states, need, x, need_x, P, G, prefix, nm, out):
        """cat(states[-nm:]) -> [conv -> BN -> ReLU -> dropout] -> += x -> LayerNorm  (node_search.py:59-68)"""
        C = self.C
        last = states[-nm:]
        nlast = need[-nm:]
        if nm != 1:
            seg = dict(W=P[prefix + '.out_conv.weight'], bias=P[prefix + '.out_conv.bias'], M=C,
                       rm=P[prefix + '.bn.running_mean'], rv=P[prefix + '.bn.running_var'],
                       nbt=P[prefix + '.bn.num_batches_tracked'],
                       gW=G.get(prefix + '.out_conv.weight'), gbias=G.get(prefix + '.out_conv.bias'))
            assert nm <= N.BMNAS_MAX_SRC
            cv = self.conv(last, [C] * nm, [seg], 1, bn=True)
            self.ln_tail(cv, x, P[prefix + '.bn.weight'], P[prefix + '.bn.bias'], G.get(prefix + '.bn.weight'),
                         G.get(prefix + '.bn.bias'), P[prefix + '.ln.weight'], P[prefix + '.ln.bias'],
                         G.get(prefix + '.ln.weight'), G.get(prefix + '.ln.bias'), prefix + '.out_dropout', out,
                         need_src=nlast, need_res=need_x)
        else:
            self.ln_cat(last, [C], x, P[prefix + '.ln.weight'], P[prefix + '.ln.bias'], G.get(prefix + '.ln.weight'),
                        G.get(prefix + '.ln.bias'), False, out, need_src=nlast, need_res=need_x)

    # ------------------------------------------------------------------ composite: searchable node cell
    def node_cell_search(self, x, y, need_x, need_y, edge_w, node_w, logits, g_edge_w, g_node_w, P, G, prefix,
                         ops, ns, nm, out, t0=None):
        """NodeCell.forward node_search.py:48-70 (prefix ends with '.node_cell').
        t0: the first inner edge mix, already written by the producer of x (chained mix, x is y)"""
        states, need = [x, y], [need_x, need_y]
        off = 0
        # the one-CTA-per-sample node kernels (small batches) also write the NEXT inner edge mix and fold its
        # input-gradient pass into their backward: one launch less per inner step in each direction
        chain_ok = CHAIN_NODE and self.B < CHAIN_NODE_MAX_B and not self.fused_mixed_ok(ops)
        t_next = t0
        for i in range(ns):
            chained_in = t_next is not None
            t = t_next if chained_in else self.buf(self.B, self.C, self.L)
            self.mix(states, edge_w, off, logits, t, gw=g_edge_w, need=list(need), dots_only=chained_in)
            s = self.buf(self.B, self.C, self.L)
            pre = f'{prefix}.node_ops.{i}'
            off += len(states)
            chain = None
            t_next = None
            if chain_ok and i + 1 < ns and len(states) <= N.BMNAS_MAX_SRC:
                t_next = self.buf(self.B, self.C, self.L)
                chain = dict(w=edge_w, w_off=off, priors=list(states), need=list(need), out=t_next)
            self.node_op(t, t, ops, P, G, (lambda k, pre=pre: f'{pre}._ops.{k}'), node_w, i * len(ops), logits, s,
                         g_gamma=g_node_w, chain=chain)
            states.append(s)
            need.append(True)
        self.node_tail(states, need, x, need_x, P, G, prefix, nm, out)

    # ------------------------------------------------------------------ composite: found node cell
    def node_cell_found(self, x, y, need_x, need_y, gene, P, G, prefix, ns, nm, out):
        """Found_NodeCell.forward node.py:45-76."""
        states, need = [x, y], [need_x, need_y]
        zero = None
        for i in range(ns):
            (n0, i0), (n1, i1) = gene.inner_edges[2 * i], gene.inner_edges[2 * i + 1]
            ins, nd = [], []
            for nme, idx in ((n0, i0), (n1, i1)):
                if nme == 'skip':
                    ins.append(states[idx])
                    nd.append(need[idx])
                else:   # 'none' == Zero: x.mul(0.)  (finite inputs)
                    if zero is None:
                        zero = self.buf(self.B, self.C, self.L, zero=True)
                    ins.append(zero)
                    nd.append(False)
            s = self.buf(self.B, self.C, self.L)
            pre = f'{prefix}.node_ops.{i}'
            self.node_op(ins[0], ins[1], [gene.inner_steps[i]], P, G, (lambda k, pre=pre: pre), None, 0, False, s,
                         need_x=nd[0], need_y=nd[1])
            states.append(s)
            need.append(True)
        self.node_tail(states, need, x, need_x, P, G, prefix, nm, out)

    # ------------------------------------------------------------------ composite: cells
    def cell_tail(self, states, need, P, G, prefix, mult, out):
        """cat(states[-m:]) -> LayerNorm([C*m, L]) -> ReLU -> view(B,-1)   (model_search.py:63-67)"""
        assert mult <= N.BMNAS_MAX_SRC
        self.ln_cat(states[-mult:], [self.C] * mult, None, P[prefix + '.ln.weight'], P[prefix + '.ln.bias'],
                    G.get(prefix + '.ln.weight'), G.get(prefix + '.ln.bias'), True, out, need_src=need[-mult:])

    def cell_search(self, feats, need_feats, alphas, g_alphas, node_arch, g_node_arch, logits, P, G, prefix,
                    steps, mult, ops, ns, nm):
        """FusionCell.forward model_search.py:50-68.  node_arch[i] = (betas_i, gammas_i)."""
        states, need = list(feats), list(need_feats)
        off = 0
        for i in range(steps):
            s_in = self.buf(self.B, self.C, self.L)
            # the node cell's first inner mix reads s_in twice (x is y): one launch writes both tensors
            t0 = self.buf(self.B, self.C, self.L) if (CHAIN_MIX and ns >= 1) else None
            chain = dict(w=node_arch[i][0], w_off=0, n=2, out=t0) if t0 is not None else None
            self.mix(states, alphas, off, logits, s_in, gw=g_alphas, need=list(need), chain=chain)
            s = self.buf(self.B, self.C, self.L)
            gb, gg = g_node_arch[i] if g_node_arch is not None else (None, None)
            self.node_cell_search(s_in, s_in, True, True, node_arch[i][0], node_arch[i][1], logits, gb, gg, P, G,
                                  f'{prefix}._step_nodes.{i}.node_cell', ops, ns, nm, s, t0=t0)
            off += len(states)
            states.append(s)
            need.append(True)
        out = self.buf(self.B, mult * self.C, self.L)
        self.cell_tail(states, need, P, G, prefix, mult, out)
        return out

    def cell_found(self, feats, need_feats, genotype, P, G, prefix, ns, nm):
        """Found_Random_FusionCell.forward model.py:133-160."""
        states, need = list(feats), list(need_feats)
        steps = len(genotype.edges) // 2
        mult = len(genotype.concat)
        zero = None
        for i in range(steps):
            hs, nd = [], []
            for nme, idx in (genotype.edges[2 * i], genotype.edges[2 * i + 1]):
                if nme == 'skip':
                    hs.append(states[idx])
                    nd.append(need[idx])
                else:
                    if zero is None:
                        zero = self.buf(self.B, self.C, self.L, zero=True)
                    hs.append(zero)
                    nd.append(False)
            s = self.buf(self.B, self.C, self.L)
            self.node_cell_found(hs[0], hs[1], nd[0], nd[1], genotype.steps[i], P, G,
                                 f'{prefix}._step_nodes.{i}.node_cell', ns, nm, s)
            states.append(s)
            need.append(True)
        out = self.buf(self.B, mult * self.C, self.L)
        self.cell_tail(states, need, P, G, prefix, mult, out)
        return out
