"""Trajectory recorders for the fused path (protocol of the reference's ``traj_recorder.py``).

The reference hands every recorder a deep copy of the FULL complex spectrum and inverse-transforms the
whole stacked trajectory at the end (``traj_recorder.py:46-55, 68-80, 118-125``). Here the state lives in
the half-spectrum layout, so a frame is materialised only at the steps the controller selects, and the
recorders of this module take it in PHYSICAL space straight from the C2R pass (``fsm_c2r``): half the
bytes per frame, no trajectory-sized ``ifftn`` afterwards. A recorder that only implements the reference
protocol (``record(step, full_spectrum_frame)``) still works: ``OperatorLike.integrate`` expands the half
spectrum for it.
"""
from typing import Callable, Optional

import torch


class IntervalController:
    """``control_func`` that fires every ``interval`` steps from ``start`` on (traj_recorder.py:7-22)."""

    def __init__(self, interval: int = 1, start: int = 0):
        self.interval, self.start = interval, start

    def __call__(self, step: int) -> bool:
        return step >= self.start and (step - self.start) % self.interval == 0


class _TrajRecorder:
    """Base protocol (traj_recorder.py:24-93): ``record`` filters by ``control_func`` and calls ``_record``."""

    #: recorders that set this accept ``record_real(step, physical_frame)`` from the hot loop
    accepts_real_frames = False

    def __init__(self, control_func: Optional[Callable[[int], bool]] = None, include_initial_state: bool = True):
        control_func = control_func if control_func is not None else (lambda step: True)
        self.control_func = control_func if include_initial_state else \
            (lambda step: False if step == 0 else control_func(step))
        self.return_in_fourier = False

    def record(self, step: int, frame: torch.Tensor):
        if self.control_func(step):
            self._record(step, frame)

    def _record(self, step: int, frame: torch.Tensor):
        raise NotImplementedError

    @staticmethod
    def _traj_ifft(trajectory: torch.Tensor) -> torch.Tensor:
        dims = tuple(-(i + 1) for i in range(trajectory.dim() - 3))
        return torch.fft.ifftn(trajectory, dim=dims)

    @property
    def trajectory(self):
        raise NotImplementedError


class AutoRecorder(_TrajRecorder):
    """Keeps the frames on the device of the simulation (traj_recorder.py:95-125).

    Frames arrive either as full spectra (reference protocol, ``record``) or as physical fields
    (``record_real``, used by ``integrate`` unless Fourier output was requested); a trajectory never mixes both.
    """
    accepts_real_frames = True

    def __init__(self, control_func: Optional[Callable[[int], bool]] = None, include_initial_state: bool = True):
        super().__init__(control_func, include_initial_state)
        self._frames = []
        self._real = None          # True: physical frames, False: spectra, None: nothing recorded yet

    def _keep(self, frame: torch.Tensor) -> torch.Tensor:
        return frame.clone()

    def _push(self, frame: torch.Tensor, real: bool, owned: bool = False):
        if self._real is None:
            self._real = real
        elif self._real != real:
            raise RuntimeError("a trajectory holds either physical frames or spectra, not both")
        self._frames.append(frame if owned else self._keep(frame))

    def _record(self, step: int, frame: torch.Tensor):
        self._push(frame, real=False)

    def record_real(self, step: int, frame: torch.Tensor):
        """``frame`` is a fresh physical field the caller hands over (no copy is taken on the device)."""
        if self.control_func(step):
            self._push(frame, real=True, owned=True)

    @property
    def trajectory(self):
        if not self._frames:
            return None
        traj = torch.stack(self._frames, dim=1)                      # (B, T, C, N...)
        if self._real:
            if self.return_in_fourier:
                dims = tuple(-(i + 1) for i in range(traj.dim() - 3))
                return torch.fft.fftn(traj, dim=dims)
            return traj
        return traj if self.return_in_fourier else self._traj_ifft(traj).real


class CPURecorder(AutoRecorder):
    """Moves every recorded frame to host memory (traj_recorder.py:127-148): long trajectories of large
    grids do not have to fit next to the state in HBM."""

    def _keep(self, frame: torch.Tensor) -> torch.Tensor:
        return frame.clone() if frame.is_cpu else frame.cpu()

    def record_real(self, step: int, frame: torch.Tensor):
        if self.control_func(step):
            self._push(frame if frame.is_cpu else frame.cpu(), real=True, owned=True)
