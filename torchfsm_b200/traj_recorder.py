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


class DiskRecorder(AutoRecorder):
    """Writes the trajectory to disk in chunks of ``cache_freq`` frames (traj_recorder.py:150-204): for trajectories
    that fit neither in HBM nor in host memory. Frames come from the C2R pass as physical fields, are staged on the host
    (``temp_cache_loc="cpu"``) or on the device (``"auto"``), and every ``cache_freq`` recorded frames one file
    ``<cache_dir>/temp_cache_<step>`` (``.npy`` with ``save_format="numpy"``) holding ``(B, T, C, N...)`` is written;
    ``trajectory`` flushes what is left and returns ``None`` as the reference does."""

    def __init__(self, control_func: Optional[Callable[[int], bool]] = None, include_initial_state: bool = True,
                 cache_dir: Optional[str] = None, cache_freq: int = 1, temp_cache_loc: str = "cpu",
                 save_format: str = "torch"):
        super().__init__(control_func, include_initial_state)
        self.cache_dir = cache_dir if cache_dir is not None else "./saved_traj/"
        self.cache_freq = int(cache_freq)
        if temp_cache_loc not in ("auto", "cpu"):
            raise ValueError("temp_cache_loc must be either 'auto' or 'cpu'.")
        if save_format not in ("numpy", "torch"):
            raise ValueError("save_format must be either 'numpy' or 'torch'.")
        self.temp_cache_loc, self.save_format = temp_cache_loc, save_format
        self.files = []

    def _keep(self, frame: torch.Tensor) -> torch.Tensor:
        return frame.cpu() if (self.temp_cache_loc == "cpu" and not frame.is_cpu) else frame.clone()

    def _push(self, frame, real, owned=False):
        if owned and self.temp_cache_loc == "cpu" and not frame.is_cpu:
            frame = frame.cpu()
        super()._push(frame, real, owned)
        self._last_step = getattr(self, "_step", 0)
        if len(self._frames) >= self.cache_freq:
            self.flush()

    def record(self, step: int, frame: torch.Tensor):
        self._step = step
        super().record(step, frame)

    def record_real(self, step: int, frame: torch.Tensor):
        self._step = step
        super().record_real(step, frame)

    def flush(self):
        if not self._frames:
            return
        import os
        chunk = AutoRecorder.trajectory.fget(self).cpu()
        os.makedirs(self.cache_dir, exist_ok=True)
        path = os.path.join(self.cache_dir, f"temp_cache_{self._last_step}")
        if self.save_format == "numpy":
            import numpy as np
            np.save(path, chunk.numpy())
            path += ".npy"
        else:
            torch.save(chunk, path)
        self.files.append(path)
        self._frames, self._real = [], None

    @property
    def trajectory(self):
        self.flush()
        return None


class RandomBatchWisedRecorder(_TrajRecorder):
    """``n_recorded_frames`` frames per SAMPLE, ``recorder_interval`` steps apart, starting at a step drawn per sample
    (traj_recorder.py:206-264; the draw is the reference's ``np.random.randint`` call, so a numpy seed gives the same
    frame indices). ``integrate`` asks ``control_func`` which steps any sample needs and fuses the steps in between;
    frames arrive as physical fields from the C2R pass. ``trajectory``: ``(B, n_recorded_frames, C, N...)``."""
    accepts_real_frames = True

    def __init__(self, simulation_steps: int, recorder_interval: int, n_recorded_frames: int = 2):
        super().__init__(None, False)
        self.simulation_steps, self.recorder_interval = simulation_steps, recorder_interval
        self.n_recorded_frames = n_recorded_frames
        self._batch_size, self._recorded_frame_id, self._trajectory, self._real = None, None, [], None
        self.control_func = self._wanted

    def prepare(self, batch_size: int):
        """Called by ``integrate`` before the first step (the reference draws at the first recorded frame)."""
        if self._batch_size is None:
            import numpy as np
            self._batch_size = batch_size
            first = np.random.randint(1, self.simulation_steps - self.recorder_interval * self.n_recorded_frames,
                                      size=batch_size)
            self._recorded_frame_id = np.stack([first + i * self.recorder_interval
                                                for i in range(self.n_recorded_frames)], axis=1)
            self._steps = set(int(s) for s in self._recorded_frame_id.reshape(-1))
            self._trajectory = [[] for _ in range(batch_size)]

    def _wanted(self, step: int) -> bool:
        return self._batch_size is None or step in self._steps

    def _take(self, step, frame, real):
        self.prepare(frame.shape[0])
        if self._real is None:
            self._real = real
        for b in range(self._batch_size):
            if step in self._recorded_frame_id[b]:
                self._trajectory[b].append(frame[b].clone())

    def record(self, step: int, frame: torch.Tensor):
        self._take(step, frame, real=False)

    def record_real(self, step: int, frame: torch.Tensor):
        self._take(step, frame, real=True)

    @property
    def trajectory(self):
        if not self._trajectory:
            return None
        trajs = torch.stack([torch.stack(t, dim=0) for t in self._trajectory], dim=0)
        self._trajectory = []
        dims = tuple(-(i + 1) for i in range(trajs.dim() - 3))
        if self._real:
            return torch.fft.fftn(trajs, dim=dims) if self.return_in_fourier else trajs
        return trajs if self.return_in_fourier else torch.fft.ifftn(trajs, dim=dims).real
