"""Minimal trajectory recorder with the reference's protocol (``traj_recorder.py:7-125``): the hot
loop calls ``record(step, full_spectrum_frame)``; ``trajectory`` stacks the frames on dim 1."""
from typing import Callable, Optional

import torch


class IntervalController:
    def __init__(self, interval: int = 1, start: int = 0):
        self.interval, self.start = interval, start

    def __call__(self, step: int) -> bool:
        return step >= self.start and (step - self.start) % self.interval == 0


class AutoRecorder:
    def __init__(self, control_func: Optional[Callable[[int], bool]] = None, include_initial_state: bool = True):
        control_func = control_func if control_func is not None else (lambda step: True)
        self.control_func = control_func if include_initial_state else \
            (lambda step: False if step == 0 else control_func(step))
        self.return_in_fourier = False
        self._frames = []

    def record(self, step: int, frame: torch.Tensor):
        if self.control_func(step):
            self._frames.append(frame.clone())

    @property
    def trajectory(self):
        if not self._frames:
            return None
        traj = torch.stack(self._frames, dim=1)
        if self.return_in_fourier:
            return traj
        dims = tuple(-(i + 1) for i in range(traj.dim() - 3))
        return torch.fft.ifftn(traj, dim=dims).real
