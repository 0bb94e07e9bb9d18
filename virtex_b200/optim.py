"""Optimiser-side mirrors: Lookahead (virtex/optim/lookahead.py:14-129) and the linear-warm-up LR schedules
(virtex/optim/lr_scheduler.py:9-183).  These are host-side torch.optim plumbing for the drop-in API; the throughput
path uses the fused device-side tail in `virtex_b200.trainer` with identical arithmetic."""
import bisect
import math
from collections import defaultdict
from typing import Any, Callable, Dict, List

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR


class Lookahead(Optimizer):
    """k fast steps of the wrapped optimiser, then `fast <- alpha*fast + (1-alpha)*slow; slow <- fast`."""

    def __init__(self, optimizer: Optimizer, k: int = 5, alpha: float = 0.8):
        self.optimizer = optimizer
        self.k = k
        self.alpha = alpha
        self._k_counter = 0
        self.state: Dict[Any, Any] = defaultdict(dict)
        self._cache_slow()

    def _cache_slow(self):
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                self.state[p]["slow_params"] = p.data.clone()

    def __getstate__(self):
        return {"state": self.state, "optimizer": self.optimizer, "alpha": self.alpha, "k": self.k,
                "_k_counter": self._k_counter}

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def zero_grad(self, *a, **k):
        self.optimizer.zero_grad(*a, **k)

    def state_dict(self):
        return self.optimizer.state_dict()  # slow weights / counter are not serialised (as in the reference)

    def load_state_dict(self, state_dict: Dict[str, Any]):
        self.optimizer.load_state_dict(state_dict)
        self._cache_slow()

    def step(self, closure: Callable = None):
        loss = self.optimizer.step(closure)
        self._k_counter += 1
        if self._k_counter >= self.k:
            self._k_counter = 0
            for group in self.optimizer.param_groups:
                for p in group["params"]:
                    slow = self.state[p]["slow_params"]
                    p.data.mul_(self.alpha).add_(slow, alpha=1.0 - self.alpha)
                    slow.copy_(p.data)
        return loss

    def load_slow_weights(self):
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                st = self.state[p]
                st["backup_params"] = p.data.clone()
                p.data.copy_(st["slow_params"])

    def restore_fast_weights(self):
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                st = self.state[p]
                p.data.copy_(st.pop("backup_params"))


# ------------------------------------------------------------------------------------------------- LR multipliers
def warmup_none(step: int, warmup_steps: int) -> float:
    return step / float(max(1, warmup_steps)) if step < warmup_steps else 1.0


def warmup_multistep(step, warmup_steps, milestones, gamma):
    if step < warmup_steps:
        return step / float(max(1, warmup_steps))
    return gamma ** bisect.bisect_right(milestones, step)


def warmup_linear(step, warmup_steps, total_steps):
    if step < warmup_steps:
        return step / float(max(1, warmup_steps))
    return max(0.0, float(total_steps - step) / float(max(1, total_steps - warmup_steps)))


def warmup_cosine(step, warmup_steps, total_steps):
    if step < warmup_steps:
        return step / float(max(1, warmup_steps))
    cos_factor = (step - warmup_steps) / (total_steps - warmup_steps)
    return max(0.0, math.cos(cos_factor * (math.pi / 2)) ** 2)


class LinearWarmupNoDecayLR(LambdaLR):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int, last_epoch: int = -1):
        assert warmup_steps < total_steps, "Warmup steps should be less than total steps."
        self.tsteps, self.wsteps = total_steps, warmup_steps
        super().__init__(optimizer, lambda s: warmup_none(s, warmup_steps), last_epoch)


class LinearWarmupMultiStepLR(LambdaLR):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int, milestones: List[int], gamma: float = 0.1,
                 last_epoch: int = -1):
        self.wsteps, self.milestones, self.gamma = warmup_steps, list(milestones), gamma
        assert self.milestones == sorted(self.milestones), "milestones must be increasing"
        assert self.milestones and self.milestones[0] > warmup_steps, "first milestone must be after warmup"
        assert self.milestones[-1] < total_steps, "last milestone must be less than total steps"
        super().__init__(optimizer, lambda s: warmup_multistep(s, warmup_steps, self.milestones, gamma), last_epoch)


class LinearWarmupLinearDecayLR(LambdaLR):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int, last_epoch: int = -1):
        assert warmup_steps < total_steps, "Warmup steps should be less than total steps."
        self.tsteps, self.wsteps = total_steps, warmup_steps
        super().__init__(optimizer, lambda s: warmup_linear(s, warmup_steps, total_steps), last_epoch)


class LinearWarmupCosineAnnealingLR(LambdaLR):
    def __init__(self, optimizer, total_steps: int, warmup_steps: int, last_epoch: int = -1):
        assert warmup_steps < total_steps, "Warmup steps should be less than total steps."
        self.tsteps, self.wsteps = total_steps, warmup_steps
        super().__init__(optimizer, lambda s: warmup_cosine(s, warmup_steps, total_steps), last_epoch)


def lr_multiplier_fn(name: str, total_steps: int, warmup_steps: int, milestones=(), gamma=0.1):
    """The same schedules as plain functions of the step (used by the fused device-side optimiser)."""
    if name == "none":
        return lambda s: warmup_none(s, warmup_steps)
    if name == "multistep":
        return lambda s: warmup_multistep(s, warmup_steps, list(milestones), gamma)
    if name == "linear":
        return lambda s: warmup_linear(s, warmup_steps, total_steps)
    if name == "cosine":
        return lambda s: warmup_cosine(s, warmup_steps, total_steps)
    raise KeyError(f"unknown LR schedule {name}")
