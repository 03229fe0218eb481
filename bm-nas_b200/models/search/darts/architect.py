"""Architecture update -- drop-in for models/search/darts/architect.py:8-29.

``Architect(model, args, criterion, optimizer).step(x, y, logger)`` = zero the arch grads,
one forward/backward of the whole model on the validation batch, one optimiser step on
alpha/beta/gamma (first order: what the reference runs).  Any optimiser object works;
``bmnas.optim.FusedAdam`` runs the update as one multi-tensor CUDA kernel.

Second order.  The reference carries the flag for the unrolled DARTS update
(main_darts_found_ntu.py:48 ``--unrolled``) but never implements it -- its Architect is the
first-order half of DARTS' architect.py.  ``step_unrolled`` supplies the other half with the
same model surface (Liu et al., DARTS, eq. 7-8; finite-difference Hessian-vector product):

    w'      = w - xi * (grad_w L_train(w, a) + wd * w)                     one virtual weight step
    d_a     = grad_a L_val(w', a) ,  d_w' = grad_w' L_val(w', a)
    w+-     = w +- eps * d_w' ,  eps = 0.01 / ||d_w'||
    grad_a  = d_a - xi * (grad_a L_train(w+, a) - grad_a L_train(w-, a)) / (2 eps)

Four forward/backward passes through the same launch plans; the weights are perturbed in
place and put back, BatchNorm running statistics and the dropout step counters are restored,
so the unrolled step leaves nothing behind but the new alpha/beta/gamma.
"""
import torch


class Architect(object):
    def __init__(self, model, args, criterion, optimizer):
        self.network_weight_decay = args.weight_decay
        self.criterion = criterion
        self.model = model
        self.optimizer = optimizer

    def log_learning_rate(self, logger):
        for group in self.optimizer.param_groups:
            logger.info("Architecture Learning Rate: {}".format(group['lr']))
            break

    def step(self, input_valid, target_valid, logger=None):
        self.optimizer.zero_grad()
        self._backward_step(input_valid, target_valid)
        self.optimizer.step()

    def _backward_step(self, input_valid, target_valid):
        loss = self.criterion(self.model(input_valid), target_valid)
        loss.backward()

    # ------------------------------------------------------------------ second order (unrolled) update
    def _weights(self):
        return [p for p in self.model.parameters() if p.requires_grad]

    def _grads(self, x, y):
        """one forward/backward; returns (weight grads, arch grads) as fresh tensors"""
        ws, arch = self._weights(), list(self.model.arch_parameters())
        for t in ws + arch:
            t.grad = None
        from bmnas import runtime as _rt
        _rt.clear_dirty(ws + arch)
        loss = self.criterion(self.model(x), y)
        loss.backward()
        return [w.grad.detach().clone() for w in ws], [a.grad.detach().clone() for a in arch]

    def step_unrolled(self, input_train, target_train, input_valid, target_valid, eta, logger=None):
        """eta: the weight learning rate xi of the virtual step (the scheduler's current value)"""
        ws, arch = self._weights(), list(self.model.arch_parameters())
        buffers = [b.detach().clone() for b in self.model.buffers()]
        saved = [w.detach().clone() for w in ws]
        self.optimizer.zero_grad()
        gw, _ = self._grads(input_train, target_train)
        with torch.no_grad():
            for w, g in zip(ws, gw):
                w.sub_(eta * (g + self.network_weight_decay * w))                 # w' (virtual step)
        dw, da = self._grads(input_valid, target_valid)
        norm = torch.sqrt(sum((g.double() ** 2).sum() for g in dw)).item()
        eps = 0.01 / max(norm, 1e-30)
        with torch.no_grad():
            for w, w0, g in zip(ws, saved, dw):
                w.copy_(w0 + eps * g)
        _, gp = self._grads(input_train, target_train)
        with torch.no_grad():
            for w, w0, g in zip(ws, saved, dw):
                w.copy_(w0 - eps * g)
        _, gn = self._grads(input_train, target_train)
        with torch.no_grad():
            for w, w0 in zip(ws, saved):
                w.copy_(w0)
            for b, b0 in zip(self.model.buffers(), buffers):
                b.copy_(b0)
        for t in ws:
            t.grad = None
        from bmnas import runtime as _rt
        _rt.clear_dirty(ws + arch)
        for a, d, p, n in zip(arch, da, gp, gn):
            g = d - eta * (p - n) / (2.0 * eps)
            if a.grad is None:
                a.grad = g
            else:
                a.grad.copy_(g)
        self.optimizer.step()
