"""Exploration-noise schedules with the interface of pql/utils/schedule_util.py: ``step()`` advances
the schedule and returns the new value, ``val()`` returns the current one.  Host-side scalars only;
the value sequences are pinned to the reference's by tests/golden/schedules.json."""
import math


class _Schedule:
    def __init__(self, start_val, total_iters):
        self.start_val = start_val
        self.total_iters = total_iters      # None: never saturates
        self.count = 0
        self.last_val = start_val

    def _next(self):
        raise NotImplementedError

    def step(self):
        saturated = self.total_iters is not None and self.count > self.total_iters
        if not saturated:
            self.last_val = self._next()
            self.count += 1
        return self.last_val

    def val(self):
        return self.last_val


class LinearSchedule(_Schedule):
    """start_val -> end_val in total_iters steps, then constant (schedule_util.py:4-23)."""

    def __init__(self, start_val, end_val, total_iters=5):
        super().__init__(start_val, total_iters)
        self.end_val = end_val

    def _next(self):
        return self.count / self.total_iters * (self.end_val - self.start_val) + self.start_val


class ExponentialSchedule(_Schedule):
    """val <- val * gamma per step until end_val would be passed (schedule_util.py:26-48)."""

    def __init__(self, start_val, gamma, end_val=None):
        iters = None if end_val is None else int((math.log(end_val) - math.log(start_val)) / math.log(gamma))
        super().__init__(start_val, iters)
        self.gamma, self.end_val = gamma, end_val

    def _next(self):
        return self.last_val * self.gamma
