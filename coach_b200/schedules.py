"""Schedules used on the hot path (PER beta), mirroring ``rl_coach/schedules.py:23-63``.

``LinearSchedule.step`` keeps the reference's *recurrence* (repeated subtraction + ``np.clip``) rather than a closed
form: the accumulated floating-point error is observable in the importance weights, so it is part of the contract
(SURVEY.md quirk Q5).
"""
import numpy as np


class Schedule(object):
    def __init__(self, initial_value: float):
        self.initial_value = initial_value
        self.current_value = initial_value

    def step(self):
        raise NotImplementedError("")


class ConstantSchedule(Schedule):
    def step(self):
        pass


class LinearSchedule(Schedule):
    def __init__(self, initial_value: float, final_value: float, decay_steps: int):
        super().__init__(initial_value)
        self.final_value = final_value
        self.decay_steps = decay_steps
        self.decay_delta = (initial_value - final_value) / float(decay_steps)

    def step(self):
        self.current_value -= self.decay_delta
        if self.final_value < self.initial_value:
            self.current_value = np.clip(self.current_value, self.final_value, self.initial_value)
        if self.final_value > self.initial_value:
            self.current_value = np.clip(self.current_value, self.initial_value, self.final_value)
