"""Learning-rate schedulers (reference: DeepFlows/optim/scheduler.py:4-59). `step()` is called once per epoch and
rewrites `optimizer.lr`, which the fused optimizer kernels read from device memory at the next step (a captured
training step picks the new value up through `CapturedStep.note_optimizer_step`). Each schedule is a rule
`(epoch, current lr) -> new lr or None`; None leaves the optimizer alone."""
import math


def _half_cosine(peak, floor, t, period):
    """floor + (peak - floor) * (1 + cos(pi t / period)) / 2"""
    return floor + (peak - floor) * (1 + math.cos(math.pi * t / period)) / 2


class LRScheduler:
    def __init__(self, optimizer) -> None:
        self.optimizer = optimizer
        self.last_epoch = -1

    def rule(self, epoch, lr):
        return None

    def step(self):
        self.last_epoch += 1
        new_lr = self.rule(self.last_epoch, getattr(self.optimizer, "lr", None))
        if new_lr is not None:
            self.optimizer.lr = new_lr


class StepLR(LRScheduler):
    """Multiplies the CURRENT lr by gamma at every positive multiple of step_size (so manual changes to
    optimizer.lr between epochs are kept, as in the reference)."""

    def __init__(self, optimizer, step_size: int, gamma: float = 0.1) -> None:
        super().__init__(optimizer)
        self.step_size, self.gamma = step_size, gamma

    def rule(self, epoch, lr):
        due = epoch != 0 and epoch % self.step_size == 0
        return lr * self.gamma if due and lr is not None else None


class CosineAnnealingLR(LRScheduler):
    """Half-cosine from the optimizer's initial lr to eta_min, restarting every T_max epochs."""

    def __init__(self, optimizer, T_max: int, eta_min: float = 0.0) -> None:
        super().__init__(optimizer)
        self.T_max, self.eta_min = T_max, eta_min
        self.base_lr = getattr(optimizer, "lr", None)

    def rule(self, epoch, lr):
        if self.base_lr is None:
            return None
        return _half_cosine(self.base_lr, self.eta_min, epoch % self.T_max, self.T_max)


class WarmupCosineLR(LRScheduler):
    """Linear ramp warmup_start_lr -> base_lr over warmup_epochs, then a half-cosine to eta_min over T_max epochs
    (the ResNet script: warmup 5, T_max = epochs, eta_min 1e-5)."""

    def __init__(self, optimizer, warmup_epochs: int, T_max: int, base_lr: float = None, warmup_start_lr: float = 0.0,
                 eta_min: float = 0.0) -> None:
        super().__init__(optimizer)
        self.warmup_epochs, self.T_max, self.eta_min = warmup_epochs, T_max, eta_min
        self.warmup_start_lr = warmup_start_lr
        self.base_lr = getattr(optimizer, "lr", None) if base_lr is None else base_lr

    def rule(self, epoch, lr):
        if self.base_lr is None:
            return None
        if 0 < self.warmup_epochs and epoch <= self.warmup_epochs:
            ramp = epoch / max(1, self.warmup_epochs)
            return self.warmup_start_lr + (self.base_lr - self.warmup_start_lr) * ramp
        return _half_cosine(self.base_lr, self.eta_min, max(0, epoch - self.warmup_epochs), max(1, self.T_max))
