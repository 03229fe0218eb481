"""FusedAdam: torch.optim.Adam semantics (coupled L2 decay, bias correction) in ONE
multi-tensor CUDA launch, with lr and step count resident on the device so that a
captured CUDA graph picks up the schedule value written before each replay.
Replaces the two Adam instances of the search (ntu_darts_searchable.py:42 and :46-47) and
the per-iteration ``optimizer.load_state_dict`` LR push (scheduler.py:42-46).
"""
import ctypes

import torch

from . import native as N


def _capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()

BLOCK_ELEMS = 1024


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.grad_scale = 1.0
        self.lr_ring = 0      # > 0 (set before the first step): the device lr is a ring of this many schedule values, see write_lr_ring
        self._g = {}          # group index -> device state

    # -------------------------------------------------------------- device state
    def _group_state(self, gi, group):
        ps = [p for p in group['params'] if p.grad is not None]
        if not ps:
            return None
        sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        st = self._g.get(gi)
        if st is not None and st['sig'] == sig:
            return st
        if _capturing():
            raise RuntimeError('FusedAdam: parameter/gradient pointers changed during CUDA-graph capture; '
                               'run a warm-up step first')
        dev = ps[0].device
        old = st
        sizes = [p.numel() for p in ps]
        padded = [(n + 3) // 4 * 4 for n in sizes]
        tot = sum(padded)
        if old is not None and old['tot'] == tot:
            m, v, step, lr = old['m'], old['v'], old['step'], old['lr']
        else:
            m = torch.zeros(tot, device=dev)
            v = torch.zeros(tot, device=dev)
            step = torch.zeros(1, dtype=torch.int64, device=dev)
            lr = torch.full((max(1, self.lr_ring),), float(group['lr']), device=dev)
        table = (N.bmnas_adam_tensor * len(ps))()
        off, blk = 0, 0
        for i, p in enumerate(ps):
            if not p.is_contiguous() or not p.grad.is_contiguous() or p.dtype != torch.float32:
                raise ValueError('FusedAdam needs contiguous fp32 parameters and gradients')
            table[i].p = p.data_ptr()
            table[i].g = p.grad.data_ptr()
            table[i].m = m.data_ptr() + off * 4
            table[i].v = v.data_ptr() + off * 4
            table[i].n = sizes[i]
            table[i].block_start = blk
            off += padded[i]
            blk += (sizes[i] + BLOCK_ELEMS - 1) // BLOCK_ELEMS
        raw = bytes(table)
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        dtab = host.to(dev)
        st = dict(sig=sig, tot=tot, m=m, v=v, step=step, lr=lr, table=dtab, n=len(ps), blocks=blk,
                  counter=torch.zeros(1, dtype=torch.int32, device=dev), lr_host=float(group['lr']),
                  params=ps)
        self._g[gi] = st
        if getattr(self, '_pending_fused', None):
            self._apply_pending()            # a checkpoint loaded before the device state existed
        return st

    def set_lr(self, lr):
        """write the learning rate for the next step: host value + device scalar (stream-ordered fill, so a
        captured step replayed afterwards sees it)"""
        for gi, group in enumerate(self.param_groups):
            group['lr'] = float(lr)
            st = self._g.get(gi)
            if st is not None:
                st['lr'].fill_(float(lr))
                st['lr_host'] = float(lr)

    def note_lr(self, lr):
        """host bookkeeping only (param_groups[i]['lr'] for loggers): the device already holds the value (write_lr_ring)"""
        for gi, group in enumerate(self.param_groups):
            group['lr'] = float(lr)
            st = self._g.get(gi)
            if st is not None:
                st['lr_host'] = float(lr)

    def write_lr_ring(self, start, values):
        """schedule values of the optimiser steps start, start + 1, ... (step = how many steps this optimiser has taken
        before the one that uses the value) go to ring[(start + i) % lr_ring]: ONE stream-ordered upload for hundreds of
        captured steps instead of a scalar write before every replay"""
        R = self.lr_ring
        assert R > 0 and len(values) <= R
        ok = True
        for gi, group in enumerate(self.param_groups):
            st = self._g.get(gi)
            if st is None:                   # no device state yet: it will be created with every entry = group['lr']
                group['lr'] = float(values[0])
                ok = False
                continue
            idx = torch.tensor([(start + i) % R for i in range(len(values))], dtype=torch.int64)
            host = torch.tensor(values, dtype=torch.float32)
            st['lr'].index_copy_(0, idx.to(st['lr'].device), host.to(st['lr'].device))
            group['lr'] = st['lr_host'] = float(values[0])
        return ok

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        s = None
        if torch.cuda.is_available() and not N.VALIDATE_ONLY:
            from .program import join_side          # weight gradients may still be in flight on the side branch
            for group in self.param_groups:
                if group['params'] and group['params'][0].is_cuda:
                    join_side(group['params'][0].device)
                    break
        for gi, group in enumerate(self.param_groups):
            st = self._group_state(gi, group)
            if st is None:
                continue
            if st['lr_host'] != float(group['lr']):
                if _capturing():
                    raise RuntimeError('FusedAdam: change the lr with set_lr() before replaying a captured step')
                st['lr'].fill_(float(group['lr']))
                st['lr_host'] = float(group['lr'])
            p = N.bmnas_adam_params()
            p.n_tensors, p.block_elems, p.total_blocks = st['n'], BLOCK_ELEMS, st['blocks']
            p.tensors = st['table'].data_ptr()
            p.lr = st['lr'].data_ptr()
            p.beta1, p.beta2 = group['betas']
            p.eps, p.weight_decay, p.grad_scale = group['eps'], group['weight_decay'], self.grad_scale
            p.step = st['step'].data_ptr()
            p.counter = st['counter'].data_ptr()
            p.lr_ring = self.lr_ring
            s = s or N.current_stream()
            N.launch('bmnas_adam_step', ctypes.byref(p), s)
        self._clear_dirty()
        return loss

    def _clear_dirty(self):
        from . import runtime as _rt
        _rt.clear_dirty([p for g in self.param_groups for p in g['params']])

    # -------------------------------------------------------------- state (checkpoint / warm-up restore)
    def state_snapshot(self):
        """clones of the device-side moments, step counters and learning rates per parameter group"""
        return {gi: {k: st[k].clone() for k in ('m', 'v', 'step', 'lr')} | {'lr_host': st['lr_host']}
                for gi, st in self._g.items()}

    def state_restore(self, snap):
        """put back a state_snapshot(); groups that did not exist at snapshot time restart from zero"""
        with torch.no_grad():
            for gi, st in self._g.items():
                old = snap.get(gi)
                if old is not None and old['m'].numel() == st['m'].numel():
                    for k in ('m', 'v', 'step', 'lr'):
                        st[k].copy_(old[k])
                    st['lr_host'] = old['lr_host']
                else:
                    st['m'].zero_(); st['v'].zero_(); st['step'].zero_()

    def state_dict(self):
        """torch.optim state_dict plus the fused moments ('bmnas_fused': per group m, v, step) -- the moments live in
        flat device buffers outside optimizer.state"""
        sd = super().state_dict()
        sd['bmnas_fused'] = {gi: {k: st[k].detach().cpu().clone() for k in ('m', 'v', 'step')} for gi, st in self._g.items()}
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        fused = sd.pop('bmnas_fused', None)
        super().load_state_dict(sd)
        self._pending_fused = fused
        self._apply_pending()

    def _apply_pending(self):
        fused = getattr(self, '_pending_fused', None)
        if not fused:
            return
        with torch.no_grad():
            for gi in list(fused):
                st = self._g.get(gi)
                if st is not None and st['m'].numel() == fused[gi]['m'].numel():
                    for k in ('m', 'v', 'step'):
                        st[k].copy_(fused[gi][k].to(st[k].device))
                    del fused[gi]

    def zero_grad(self, set_to_none=True):
        """Gradients live in a static arena; the next backward into these parameters starts from zero again (the
        arena's dirty marks are dropped -- no memset needed, every plan zeroes its span before it accumulates)."""
        self._clear_dirty()
        return None
