"""Per-batch learning-rate schedules of the reference (host side, fp64).

Same class names, constructor arguments and observable behaviour as
/root/reference/models/auxiliary/scheduler.py:12-62, so drivers that build them by name keep
working.  ``update_optimizer`` writes the LR straight into ``optimizer.param_groups`` instead of the
reference's state_dict()/load_state_dict() round trip (which deep-copies the Adam moments every
batch) -- the visible effect, ``group['lr'] == eta``, is the same.
"""
import math


class LRCosineAnnealingScheduler:
    """Cosine annealing with warm restarts, evaluated once per batch.

    Reference quirk kept on purpose (scheduler.py:29-40): the period position is the float
    ``iteration / num_batches_per_epoch`` and a restart fires only when eta comes within 1e-10 of
    eta_min, i.e. when that position lands on Ti exactly; with a non-integral batches-per-epoch
    the cosine just keeps oscillating with period 2*Ti.
    """

    def __init__(self, eta_max, eta_min, Ti, Tmultiplier, num_batches_per_epoch):
        self.eta_max = eta_max
        self.eta_min = eta_min
        self.Ti = Ti
        self.Tm = Tmultiplier
        self.nbpe = num_batches_per_epoch
        self.Tcur = 0.0
        self.iteration_counter = 0.0
        self.eta = eta_max

    def step(self):
        self.Tcur = self.iteration_counter / self.nbpe
        self.iteration_counter += 1.0
        span = self.eta_max - self.eta_min
        self.eta = self.eta_min + 0.5 * span * (1 + math.cos(math.pi * self.Tcur / self.Ti))
        eta = self.eta
        if eta <= self.eta_min + 1e-10:          # warm restart: next period is Tm times longer
            self.Tcur = 0
            self.iteration_counter = 0
            self.Ti = self.Ti * self.Tm
        return eta

    def update_optimizer(self, optimizer):
        for group in optimizer.param_groups:
            group['lr'] = self.eta


class FixedScheduler:
    def __init__(self, lr):
        self.lr = lr

    def step(self):
        return self.lr

    def update_optimizer(self, optimizer):
        for group in optimizer.param_groups:
            group['lr'] = self.lr


def is_per_batch_cosine(scheduler) -> bool:
    """True for this class and for the reference's own LRCosineAnnealingScheduler instances."""
    return isinstance(scheduler, LRCosineAnnealingScheduler) or (
        type(scheduler).__name__ == "LRCosineAnnealingScheduler" and hasattr(scheduler, "nbpe"))
