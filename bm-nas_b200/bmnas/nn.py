"""Head of the search network: classifier GEMM and loss on the same C-ABI kernels.

* Linear            -- nn.Linear drop-in (ntu_darts_searchable.py:100-101, central_classifier);
                       bmnas_linear_fwd / bmnas_linear_bwd (csrc/linear.cu: three skinny single-launch GEMMs);
                       shapes those cannot take fall back to the generic conv GEMM kernels.
* CrossEntropyLoss / BCEWithLogitsLoss -- mean-reduced criteria of the search scripts
                       (ntu_darts_searchable.py:25, mmimdb_darts_searchable.py:22), fused
                       softmax/sigmoid + gradient kernel.
* SearchHead        -- Searchable_*_Net minus backbones and reshape layers
                       (ntu_darts_searchable.py:71-179): ``fusion_net`` + ``central_classifier``
                       with the reference's attribute names, one joint gradient arena.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import native as N
from . import runtime as _rt
from . import program as _prog


# classifier + criterion + their backward as ONE kernel (bmnas_head_fused, csrc/head_fused.cu) wherever SearchHead.loss_fused
# is used (SearchStep does); 0 = the five-launch chain Linear -> criterion -> backward of both
FUSED_HEAD = os.environ.get('BMNAS_FUSED_HEAD', '1') != '0'
# SearchStep calls loss.backward() itself: the incoming gradient of the loss is the constant 1 and is not multiplied in
UNIT_LOSS_GRAD = [False]


def _conv_struct(B, L, K, M, w_fold=1):
    st = N.bmnas_conv_params()
    st.B, st.L, st.K, st.M, st.w_fold, st.n_src, st.n_seg = B, L, K, M, w_fold, 1, 1
    st.src_C[0] = K
    st.seg_M[0] = M
    return st


def _p(t):
    return None if t is None else t.data_ptr()


def _lin_ok(x, weight):
    return x.shape[1] % 4 == 0 and x.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0 and weight.shape[0] <= 1024


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gw_view, gb_view, arena=None):
        B, Kc = x.shape
        Nc = weight.shape[0]
        s = N.current_stream()
        out = torch.empty(B, Nc, device=x.device, dtype=torch.float32)
        if _lin_ok(x, weight):
            st = N.bmnas_linear_params()
            st.B, st.K, st.N = B, Kc, Nc
            st.x, st.W, st.bias, st.out = x.data_ptr(), weight.data_ptr(), _p(bias), out.data_ptr()
            N.launch('bmnas_linear_fwd', ctypes.byref(st), s)
        else:   # odd shapes: generic GEMM kernels (NT GEMM = bmnas_conv_wgrad on a bias-initialised output)
            N.launch('bmnas_bias_rows', ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(_p(bias)), B, Nc, s)
            st = _conv_struct(1, Kc, Nc, B)       # out[m=b][k=class] += sum_l x[b][l] * W[class][l]
            st.GV = x.data_ptr()
            st.src[0] = weight.data_ptr()
            st.gW[0] = out.data_ptr()
            N.launch('bmnas_conv_wgrad', ctypes.byref(st), s)
        ctx.save_for_backward(x, weight)
        ctx.views = (gw_view, gb_view)
        ctx.has_bias = bias is not None
        # autograd accumulate semantics outside runtime.STATIC_IO: parameters whose arena view already holds an unconsumed
        # gradient get it added back after this backward overwrote the view
        ctx.pending = []
        if arena is not None and not _rt.STATIC_IO[0]:
            for t, v in ((weight, gw_view), (bias, gb_view)):
                if t is not None and v is not None:
                    if id(t) in arena.dirty and t.grad is not None and t.grad.data_ptr() == v.data_ptr():
                        ctx.pending.append(v)
                    arena.dirty.add(id(t))
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        gw_view, gb_view = ctx.views
        g = g.contiguous()
        if ctx.pending:
            saved = [v.clone() for v in ctx.pending]
            res = _LinearFn._backward(ctx, g, x, weight, gw_view, gb_view)
            _prog.join_side(x.device)
            for v, old in zip(ctx.pending, saved):
                v.add_(old)
            return res
        return _LinearFn._backward(ctx, g, x, weight, gw_view, gb_view)

    @staticmethod
    def _backward(ctx, g, x, weight, gw_view, gb_view):
        B, Kc = x.shape
        Nc = weight.shape[0]
        s = N.current_stream()
        gx = None
        if _lin_ok(x, weight) and (gw_view is None or gw_view.data_ptr() % 16 == 0):
            def block():
                st = N.bmnas_linear_params()
                st.B, st.K, st.N = B, Kc, Nc
                st.x, st.W, st.gout = x.data_ptr(), weight.data_ptr(), g.data_ptr()
                return st
            want_w = gw_view is not None or gb_view is not None
            if ctx.needs_input_grad[0]:
                # the input gradient is what the rest of the backward waits for: main chain
                gx = torch.empty_like(x)
                st = block()
                st.gx = gx.data_ptr()
                if want_w and not _prog.SIDE_WGRAD:
                    st.gW, st.gbias = _p(gw_view), _p(gb_view)
                    want_w = False
                N.launch('bmnas_linear_bwd', ctypes.byref(st), s)
            if want_w:
                # dW / db only feed the optimiser: side branch (joined by the launch plan's backward, the
                # all-reduce and FusedAdam.step through program.join_side)
                st = block()
                st.gW, st.gbias = _p(gw_view), _p(gb_view)
                if _prog.SIDE_WGRAD and not N.VALIDATE_ONLY and ctx.needs_input_grad[0]:
                    main = torch.cuda.current_stream()
                    side = _prog.side_stream(x.device)
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                    g.record_stream(side)
                    N.launch('bmnas_linear_bwd', ctypes.byref(st), ctypes.c_void_p(side.cuda_stream))
                else:
                    N.launch('bmnas_linear_bwd', ctypes.byref(st), s)
            return gx, None, None, None, None, None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            st = _conv_struct(1, Kc, Nc, B)   # gx[b][l] = sum_class g[b][class] * W[class][l]
            st.W[0] = g.data_ptr()
            st.src[0] = weight.data_ptr()
            st.Z = gx.data_ptr()
            N.launch('bmnas_conv_fwd', ctypes.byref(st), s)
        if gw_view is not None:
            st = _conv_struct(1, Kc, Nc, B)   # gW[class][l] = sum_b g[b][class] * x[b][l]
            st.W[0] = g.data_ptr()
            st.GV = x.data_ptr()
            st.gsrc[0] = gw_view.data_ptr()
            N.launch('bmnas_conv_dgrad', ctypes.byref(st), s)
        if gb_view is not None:
            N.launch('bmnas_colsum', ctypes.c_void_p(gb_view.data_ptr()), ctypes.c_void_p(g.data_ptr()), B, Nc, s)
        return gx, None, None, None, None, None


def _assign_grad(t, view):
    if view is None:
        return
    if t.grad is None:
        t.grad = view
    elif t.grad.data_ptr() != view.data_ptr():
        t.grad.add_(view)


class Linear(nn.Linear):
    """nn.Linear whose forward/backward run on the bmnas GEMM kernels; weight/bias gradients are written
    straight into the module's gradient arena (see runtime.GradArena)."""

    def forward(self, x):
        if not x.is_cuda and not N.VALIDATE_ONLY:
            raise N.NativeError('bmnas.nn.Linear has no CPU implementation')
        x = x if x.is_contiguous() else x.contiguous()
        all_leaves = [p for p in (self.weight, self.bias) if p is not None and p.requires_grad]
        leaves = _rt.filter_leaves(all_leaves)       # runtime.GRAD_MODE 'arch': the classifier's dW / db are not produced
        gw = gb = ar = None
        if leaves and torch.is_grad_enabled():
            ar = _rt.arena_for(self, all_leaves, x.device)
            gw = ar.view(self.weight) if self.weight.requires_grad else None
            gb = ar.view(self.bias) if (self.bias is not None and self.bias.requires_grad) else None
        out = _LinearFn.apply(x, self.weight, self.bias, gw, gb, ar)
        if gw is not None or gb is not None:
            w, b = self.weight, self.bias

            def hook(_g, w=w, b=b, gw=gw, gb=gb):
                _assign_grad(w, gw)
                if b is not None:
                    _assign_grad(b, gb)
            if out.requires_grad:
                out.register_hook(hook)
        return out


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, kind, ws):
        logits = logits if logits.is_contiguous() else logits.contiguous()
        B, n = logits.shape
        st = N.bmnas_loss_params()
        st.B, st.n_classes, st.kind = B, n, kind
        loss = torch.empty((), device=logits.device, dtype=torch.float32)
        gl = torch.empty_like(logits)
        st.logits, st.loss, st.glogits = logits.data_ptr(), loss.data_ptr(), gl.data_ptr()
        if kind == 0:
            st.labels = target.data_ptr()
        else:
            st.targets = target.data_ptr()
        st.partials, st.counter = ws[0].data_ptr(), ws[1].data_ptr()
        N.launch('bmnas_loss_fwd', ctypes.byref(st), N.current_stream())
        ctx.save_for_backward(gl)
        ctx.shape = (B, n)
        return loss

    @staticmethod
    def backward(ctx, g):
        gl, = ctx.saved_tensors
        B, n = ctx.shape
        st = N.bmnas_loss_params()
        st.B, st.n_classes = B, n
        out = torch.empty_like(gl)
        g = g.contiguous()
        st.glogits, st.gscale, st.gout_logits = gl.data_ptr(), g.data_ptr(), out.data_ptr()
        N.launch('bmnas_loss_bwd', ctypes.byref(st), N.current_stream())
        return out, None, None, None


class _Loss(nn.Module):
    _kind = 0

    def _ws(self, device):
        ws = self.__dict__.get('_bm_ws')
        if ws is None or ws[0].device != device:
            st = N.bmnas_loss_params()
            n = int(N.lib().bmnas_loss_partials_size(ctypes.byref(st)))
            ws = (torch.zeros(n, device=device), torch.zeros(1, dtype=torch.int32, device=device))
            self.__dict__['_bm_ws'] = ws
        return ws

    def forward(self, logits, target):
        if not logits.is_cuda and not N.VALIDATE_ONLY:
            raise N.NativeError('bmnas losses have no CPU implementation')
        if self._kind == 0:
            target = target.to(torch.int64).contiguous()
        else:
            target = target.to(torch.float32).contiguous()
        return _LossFn.apply(logits, target, self._kind, self._ws(logits.device))


class CrossEntropyLoss(_Loss):
    _kind = 0


class BCEWithLogitsLoss(_Loss):
    _kind = 1


class _HeadLossFn(torch.autograd.Function):
    """loss, logits = criterion(x W^T + b, target) with d loss / d x computed by the same launch (bmnas_head_fused);
    backward only forks the classifier's dW / db onto the side branch and hands the stored input gradient on."""

    @staticmethod
    def forward(ctx, x, weight, bias, target, kind, ws, gw_view, gb_view):
        B, Kc = x.shape
        Nc = weight.shape[0]
        want_gx = bool(ctx.needs_input_grad[0])      # Function.forward runs with grad mode off: ask autograd
        logits = torch.empty(B, Nc, device=x.device, dtype=torch.float32)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        gl = torch.empty(B, Nc, device=x.device, dtype=torch.float32) if want_gx else None
        gx = torch.empty_like(x) if want_gx else None
        st = N.bmnas_head_params()
        st.B, st.K, st.N, st.kind = B, Kc, Nc, kind
        st.x, st.W, st.bias = x.data_ptr(), weight.data_ptr(), _p(bias)
        if kind == 0:
            st.labels = target.data_ptr()
        else:
            st.targets = target.data_ptr()
        st.logits, st.loss, st.glogits, st.gx = logits.data_ptr(), loss.data_ptr(), _p(gl), _p(gx)
        st.partials, st.counter = ws[0].data_ptr(), ws[1].data_ptr()
        N.launch('bmnas_head_fused', ctypes.byref(st), N.current_stream())
        ctx.x, ctx.gl, ctx.gx = x, gl, gx
        ctx.views = (gw_view, gb_view)
        ctx.dims = (B, Kc, Nc, weight.data_ptr())
        ctx.mark_non_differentiable(logits)
        ctx.set_materialize_grads(False)     # no zero tensor for the logits' (non-existent) gradient
        return loss, logits

    @staticmethod
    def backward(ctx, g, _unused):
        x, gl, gx = ctx.x, ctx.gl, ctx.gx
        gw_view, gb_view = ctx.views
        B, Kc, Nc, wptr = ctx.dims
        if not UNIT_LOSS_GRAD[0]:             # a caller that scales the loss: d/dx and d/dlogits are linear in it
            gl = gl * g
            gx = gx * g
        if gw_view is not None or gb_view is not None:
            st = N.bmnas_linear_params()
            st.B, st.K, st.N = B, Kc, Nc
            st.x, st.W, st.gout = x.data_ptr(), wptr, gl.data_ptr()
            st.gW, st.gbias = _p(gw_view), _p(gb_view)
            if _prog.SIDE_WGRAD and not N.VALIDATE_ONLY:
                # dW / db only feed the optimiser: side branch (joined by the launch plan's backward, the all-reduce and
                # FusedAdam.step through program.join_side)
                main = torch.cuda.current_stream()
                side = _prog.side_stream(x.device)
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                gl.record_stream(side)
                N.launch('bmnas_linear_bwd', ctypes.byref(st), ctypes.c_void_p(side.cuda_stream))
            else:
                N.launch('bmnas_linear_bwd', ctypes.byref(st), N.current_stream())
        return gx, None, None, None, None, None, None, None


class SearchHead(nn.Module):
    """fusion network + classifier: what the search step optimises once backbone features are given.
    ``genotype=None`` builds the searchable hypernet, otherwise the found network."""

    def __init__(self, args, num_outputs, criterion=None, genotype=None):
        super().__init__()
        from models.search.darts.model_search import FusionNetwork
        from models.search.darts.model import Found_FusionNetwork
        self.args = args
        self._criterion = criterion
        if genotype is None:
            self.fusion_net = FusionNetwork(steps=args.steps, multiplier=args.multiplier,
                                            num_input_nodes=args.num_input_nodes, num_keep_edges=2, args=args,
                                            criterion=criterion)
            mult = args.multiplier
        else:
            self.fusion_net = Found_FusionNetwork(len(genotype.edges) // 2, len(genotype.concat),
                                                  args.num_input_nodes, 2, args, criterion, genotype)
            mult = len(genotype.concat)
        self.central_classifier = Linear(args.C * args.L * mult, num_outputs)

    def _joint_arena(self, device):
        ar = self.__dict__.get('_bm_joint')
        # layout [alpha, beta, gamma | fusion weights | classifier]: the architecture gradients and the weight
        # gradients are each one contiguous span (one small and one large NCCL bucket, see search.py)
        leaves = (self.arch_parameters() + [p for p in self.fusion_net.parameters() if p.requires_grad] +
                  [p for p in self.central_classifier.parameters() if p.requires_grad])
        if ar is None or ar.flat.device != device or not ar.covers(leaves):
            ar = _rt.GradArena(leaves, device)
            self.__dict__['_bm_joint'] = ar
            self.fusion_net._bm_arena = ar
            self.central_classifier._bm_arena = ar
            self.fusion_net.__dict__.pop('_bm_cache', None)
        return ar

    def forward(self, feats):
        self._joint_arena(feats[0].device)
        return self.central_classifier(self.fusion_net(feats))

    def loss_fused(self, feats, labels, criterion):
        """criterion(self(feats), labels) with the classifier, the criterion and the gradient they send back into the fusion
        network in ONE launch (bmnas_head_fused).  Returns (loss, logits), or None when the fused kernel does not take the
        case (then call criterion(self(feats), labels)): it needs a bmnas criterion, the overwrite semantics of
        runtime.static_io (or no gradient at all) and a shape bmnas_head_supported accepts."""
        if not FUSED_HEAD or not isinstance(criterion, _Loss):
            return None
        if torch.is_grad_enabled() and not _rt.STATIC_IO[0]:
            return None
        lin = self.central_classifier
        dev = feats[0].device
        if not feats[0].is_cuda and not N.VALIDATE_ONLY:
            raise N.NativeError('bmnas.nn.SearchHead has no CPU implementation')
        kind = criterion._kind
        st = N.bmnas_head_params()
        st.B, st.K, st.N, st.kind = feats[0].shape[0], lin.in_features, lin.out_features, kind
        # pointer fields only need to be non-NULL / aligned for the support check
        st.x = st.W = st.logits = st.loss = st.partials = st.counter = 16
        st.labels = st.targets = 16
        if not N.lib().bmnas_head_supported(ctypes.byref(st)) or lin.weight.data_ptr() % 16:
            return None
        ws = self.__dict__.get('_bm_head_ws')
        n = int(N.lib().bmnas_head_partials_size(ctypes.byref(st)))
        if ws is None or ws[0].device != dev or ws[0].numel() < n:
            ws = (torch.zeros(max(n, 1), device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
            self.__dict__['_bm_head_ws'] = ws
        self._joint_arena(dev)
        x = self.fusion_net(feats)
        x = x if x.is_contiguous() else x.contiguous()
        if x.data_ptr() % 16:
            return None
        target = labels.to(torch.int64).contiguous() if kind == 0 else labels.to(torch.float32).contiguous()
        all_leaves = [p for p in (lin.weight, lin.bias) if p is not None and p.requires_grad]
        leaves = _rt.filter_leaves(all_leaves)
        gw = gb = None
        if leaves and torch.is_grad_enabled():
            ar = _rt.arena_for(lin, all_leaves, dev)
            gw = ar.view(lin.weight) if lin.weight.requires_grad else None
            gb = ar.view(lin.bias) if (lin.bias is not None and lin.bias.requires_grad) else None
        loss, logits = _HeadLossFn.apply(x, lin.weight, lin.bias, target, kind, ws, gw, gb)
        if (gw is not None or gb is not None) and loss.requires_grad:
            w, b = lin.weight, lin.bias

            def hook(_g, w=w, b=b, gw=gw, gb=gb):
                _assign_grad(w, gw)
                if b is not None:
                    _assign_grad(b, gb)
            loss.register_hook(hook)
        return loss, logits

    def genotype(self):
        return self.fusion_net.genotype()

    def arch_parameters(self):
        return self.fusion_net.arch_parameters() if hasattr(self.fusion_net, 'arch_parameters') else []

    def central_params(self):
        return [{'params': self.fusion_net.parameters()}, {'params': self.central_classifier.parameters()}]

    def _loss(self, feats, labels):
        return self._criterion(self(feats), labels)
