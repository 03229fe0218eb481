"""First-order architecture update -- drop-in for models/search/darts/architect.py:8-29.

``Architect(model, args, criterion, optimizer).step(x, y, logger)`` = zero the arch grads,
one forward/backward of the whole model on the validation batch, one optimiser step on
alpha/beta/gamma.  Any optimiser object works; ``bmnas.optim.FusedAdam`` runs the update
as one multi-tensor CUDA kernel.
"""


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
