"""Linear ramp used by the slotted RSSM's mixer coefficient (reference: rl_sandbox/utils/schedulers.py:9-25)."""
import numpy as np


class Scheduler:
    def step(self) -> float:
        raise NotImplementedError


class LinearScheduler(Scheduler):
    """value(t) ramps from `initial_value` at t=0 to `final_value` at t=duration-1 and stays there."""

    def __init__(self, initial_value, final_value, duration):
        self._init, self._final = initial_value, final_value
        self._dur = duration - 1
        self._curr_t = 0

    @property
    def val(self) -> float:
        if self._curr_t >= self._dur:
            return self._final
        return float(np.interp(self._curr_t, [0, self._dur], [self._init, self._final]))

    def step(self) -> float:
        current = self.val
        self._curr_t += 1
        return current
