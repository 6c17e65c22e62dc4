# -*- coding: utf-8 -*-
"""Step-size schedules for the host-side theta optimisers (Adam / SGD run between kernel launches).

API kept from the reference so that user code carries over: a schedule is an object with ``get(t)``;
``ConstantLearningRate(lr)`` and ``ExponentialLearningRate(lr_start, lr_end, steps)`` are the two the experiments
use (interface: pypsmf/psmf/learning_rate.py).  The exponential schedule interpolates geometrically from
``lr_start`` at t = 0 to ``lr_end`` at t = ``steps`` and keeps decaying beyond it."""

from __future__ import annotations

import math
from typing import Protocol, runtime_checkable


@runtime_checkable
class BaseLearningRate(Protocol):
    """Anything with ``get(t) -> float``; subclass it or just provide the method."""

    def get(self, t) -> float: ...


class ConstantLearningRate:
    __slots__ = ("lr",)

    def __init__(self, lr):
        if not math.isfinite(lr):
            raise ValueError("lr must be finite")
        self.lr = float(lr)

    def get(self, t) -> float:
        return self.lr

    def __repr__(self):
        return "ConstantLearningRate(%g)" % self.lr


class ExponentialLearningRate:
    __slots__ = ("lr_start", "lr_end", "steps", "_ratio")

    def __init__(self, lr_start, lr_end, steps):
        if lr_start <= 0 or lr_end <= 0:
            raise ValueError("lr_start and lr_end must be positive")
        if steps <= 0:
            raise ValueError("steps must be positive")
        self.lr_start, self.lr_end, self.steps = float(lr_start), float(lr_end), steps
        self._ratio = self.lr_end / self.lr_start

    def get(self, t) -> float:
        return self.lr_start * self._ratio ** (t / self.steps)

    def __repr__(self):
        return "ExponentialLearningRate(%g -> %g over %s steps)" % (self.lr_start, self.lr_end, self.steps)
