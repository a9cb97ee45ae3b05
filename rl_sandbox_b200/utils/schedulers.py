"""Linear ramp used by the slotted RSSM's mixer coefficient (reference: rl_sandbox/utils/schedulers.py:9-25)."""
import numpy as np


class Scheduler:
    def step(self):
        raise NotImplementedError


class LinearScheduler(Scheduler):
    def __init__(self, initial, final, iters):
        self._iters = max(1, iters - 1)
        self._val = initial
        self._initial, self._final = initial, final
        self._curr = 0

    @property
    def val(self):
        return float(np.interp(self._curr, [0, self._iters], [self._initial, self._final]))

    def step(self):
        v = self.val
        self._curr += 1
        return v
