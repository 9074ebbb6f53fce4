"""Mesh and spectral tables (host-side mirror of the reference's ``torchfsm/mesh.py``).

Everything here is set-up work done with torch ops on the target device; the hot path only
consumes the small 1-D tables and the coefficient tables derived from them. The formulas
follow the reference exactly (same expressions, same dtype, same order of operations) because
the ETD coefficient tables built from them are inputs to the CUDA path:
  per-axis frequencies  mesh.py:178-192   symbols (2*pi*i*f)^n  mesh.py:399-404
  Laplacian             mesh.py:406-426   2/3-rule mask         mesh.py:443-461
"""
from typing import Optional, Sequence, Union

import torch


class _PerAxis:
    """Per-axis tables reachable both as the reference spells it (``f_mesh.bf[i]``, ``len(f_mesh.bf)``; mesh.py:148-292)
    and as a call (``f_mesh.bf(i)``, the spelling used inside this package)."""

    def __init__(self, get, n_dim, names):
        self._get, self._n_dim, self._names = get, n_dim, names

    def __len__(self):
        return self._n_dim

    def __getitem__(self, i):
        if i >= self._n_dim:
            raise ValueError(f"{self._names[i]} fft frequency is not defined" if i <= 2 else
                             f"fft frequency with id{i} is not defined")
        return self._get(i)

    __call__ = __getitem__


class MeshGrid:
    """Periodic box ``[(start, end, n_points), ...]`` (mirror of mesh.py:9-160)."""

    def __init__(self, mesh_info: Sequence[tuple], device=None, dtype=None):
        for m in mesh_info:
            if len(m) != 3:
                raise ValueError("each dimension should be a tuple of (start,end,n_points)")
        self.mesh_info = [tuple(m) for m in mesh_info]
        self.n_dim = len(self.mesh_info)
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.dtype = dtype if dtype is not None else torch.get_default_dtype()

    @property
    def meshs(self):
        """The per-axis coordinate vectors (mesh.py:33-55 caches them under this name)."""
        return [self[i] for i in range(self.n_dim)]

    def __len__(self):
        return self.n_dim

    def __getitem__(self, i):
        # mesh.py:56-62: coordinates start at 0 whatever `start` is (kept as the reference has it, so that forcing
        # and initial conditions built from bc_mesh_grid() match for the same user script)
        a, b, n = self.mesh_info[i]
        return (b - a) * torch.arange(0, n, device=self.device, dtype=self.dtype) / n

    x = property(lambda self: self[0])
    y = property(lambda self: self[1])
    z = property(lambda self: self[2])

    def mesh_grid(self):
        g = torch.meshgrid(*[self[i] for i in range(self.n_dim)], indexing="ij")
        return g[0] if len(g) == 1 else g

    def bc_mesh_grid(self, batch_size: int = 1, n_channels: int = 1):
        g = self.mesh_grid()
        if isinstance(g, torch.Tensor):
            return g.unsqueeze(0).unsqueeze(0).repeat(batch_size, n_channels, *([1] * self.n_dim))
        return tuple(t.unsqueeze(0).unsqueeze(0).repeat(batch_size, n_channels, *([1] * self.n_dim)) for t in g)

    def to(self, device=None, dtype=None):
        self.device = torch.device(device) if device is not None else self.device
        self.dtype = dtype if dtype is not None else self.dtype


class FourierMesh:
    """Spectral tables of a periodic box in the reference's full ``(1, 1, N...)`` layout."""

    def __init__(self, mesh: Union[Sequence[tuple], MeshGrid, "FourierMesh"], device=None, dtype=None):
        if isinstance(mesh, (MeshGrid, FourierMesh)):
            self.mesh_info = mesh.mesh_info
            device = mesh.device if device is None else device
            dtype = mesh.dtype if dtype is None else dtype
        else:
            self.mesh_info = [tuple(m) for m in mesh]
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype if dtype is not None else torch.get_default_dtype()
        if self.dtype.is_complex:
            self.dtype = torch.float32 if self.dtype == torch.complex64 else torch.float64
        self.n_dim = len(self.mesh_info)
        self.shape = tuple(m[2] for m in self.mesh_info)
        self._f = [torch.fft.fftfreq(n, (b - a) / n, device=self.device, dtype=self.dtype)
                   for (a, b, n) in self.mesh_info]
        self.fft_dim = tuple(-(i + 1) for i in range(self.n_dim))
        self.f = _PerAxis(lambda i: self._f[i], self.n_dim, ("f_x", "f_y", "f_z"))
        self.bf = _PerAxis(self._bf, self.n_dim, ("bf_x", "bf_y", "bf_z"))
        self._default_rel_freq_threshold = 2 / 3
        self._plans = {}

    @property
    def cdtype(self):
        return torch.complex64 if self.dtype == torch.float32 else torch.complex128

    def _bf(self, i):
        shape = [1] * (self.n_dim + 2)
        shape[i + 2] = self.shape[i]
        return self._f[i].reshape(shape)

    f_x = property(lambda self: self.f[0])
    f_y = property(lambda self: self.f[1])
    f_z = property(lambda self: self.f[2])
    bf_x = property(lambda self: self.bf[0])
    bf_y = property(lambda self: self.bf[1])
    bf_z = property(lambda self: self.bf[2])

    @property
    def bf_vector(self):
        """(1, d, N...) stacked broadcast frequencies (mesh.py:256-266)."""
        return torch.cat([self.bf[i].expand(1, 1, *self.shape) for i in range(self.n_dim)], dim=1)

    def set_default_rel_freq_threshold(self, threshold: float):
        self._default_rel_freq_threshold = threshold

    def to(self, device=None, dtype=None):
        self.__init__(self.mesh_info, device=device if device is not None else self.device,
                      dtype=dtype if dtype is not None else self.dtype)

    def grad(self, dim_i: int, order: int):
        return (2j * torch.pi * self.bf(dim_i)) ** order

    def nabla(self, order: int = 1):
        return sum(self.grad(i, order) for i in range(self.n_dim))

    def laplacian(self):
        return self.nabla(2)

    def invert_laplacian(self):
        lap = self.laplacian()                                      # mesh.py:413-419
        return torch.where(lap == 0, 1.0, 1 / lap)

    def invert_nabla(self, order: int = 1):
        nab = self.nabla(order)                                     # mesh.py:428-434
        return torch.where(nab == 0, 1.0, 1 / nab)

    def nabla_vector(self, order: int):
        return (2j * torch.pi * self.bf_vector) ** order            # mesh.py:436-441

    def low_pass_filter(self, rel_freq_threshold: Optional[float] = None) -> torch.Tensor:
        """(1, 1, N...) 0/1 mask: |f_i| <= rate * max|f_i| on every axis (mesh.py:443-461)."""
        rate = self._default_rel_freq_threshold if rel_freq_threshold is None else rel_freq_threshold
        mask = torch.ones((1, 1) + self.shape, device=self.device, dtype=self.dtype)
        for i in range(self.n_dim):
            abs_f = self.bf[i].abs()
            mask = mask * torch.where(abs_f > abs_f.max() * rate, 0, 1)
        return mask.to(device=self.device, dtype=self.dtype)

    def abs_low_pass_filter(self, abs_freq_threshold: int) -> torch.Tensor:
        mask = torch.ones((1, 1) + self.shape, device=self.device, dtype=self.dtype)     # mesh.py:463-479
        for i in range(self.n_dim):
            mask = mask * torch.where(self.bf[i].abs() > abs_freq_threshold, 0, 1)
        return mask.to(device=self.device, dtype=self.dtype)

    # ---- the transform choke point (mesh.py:481-491) on the library's passes -------------------------------
    def _plan(self, n_fields: int):
        """Transform-only plan for ``n_fields`` scalar fields (leading axes of the argument flattened)."""
        st = self._plans.get(n_fields)
        if st is None:
            from . import _cabi
            from .operator import FusedStepper
            if len(self._plans) > 4:
                self._plans.clear()
            st = self._plans[n_fields] = FusedStepper(self, n_fields, 1, _cabi.PROG_LINEAR, "RK4", 1.0, None, 0.0, None,
                                                      [n // 2 for n in self.shape], True, {})
        return st

    def fft(self, u: torch.Tensor) -> torch.Tensor:
        """Full complex spectrum of ``u`` over the last ``n_dim`` axes (``torch.fft.fftn`` in the reference): R2C passes
        of the library plus the Hermitian extension; a complex argument is two real transforms."""
        if u.is_complex():
            return self.fft(u.real) + 1j * self.fft(u.imag)
        lead = u.shape[:u.dim() - self.n_dim]
        st = self._plan(max(1, int(torch.Size(lead).numel())))
        full = st.half_to_full(st.r2c(u.reshape(-1, 1, *self.shape)))
        return full.reshape(*lead, *self.shape)

    def ifft(self, u_fft: torch.Tensor) -> torch.Tensor:
        """Complex inverse transform of ANY full spectrum (``torch.fft.ifftn``): the Hermitian part of F gives the real
        part and the Hermitian part of -iF the imaginary part, each one C2R chain of the library."""
        lead = u_fft.shape[:u_fft.dim() - self.n_dim]
        st = self._plan(max(1, int(torch.Size(lead).numel())))
        F = u_fft.to(self.cdtype).reshape(-1, 1, *self.shape)
        re = st.c2r(st.full_to_half(F))
        im = st.c2r(st.full_to_half(-1j * F))
        return torch.complex(re, im).reshape(*lead, *self.shape)

    def low_pass_kmax(self, rel_freq_threshold: float):
        """Per-axis largest kept |mode index| of the reference's low-pass mask (mesh.py:443-461)."""
        kmax = []
        for i, n in enumerate(self.shape):
            abs_f = self._f[i].abs()
            kept = ~(abs_f > abs_f.max() * rel_freq_threshold)
            idx = torch.arange(n, device=self.device)
            m = torch.where(idx <= n // 2, idx, n - idx)  # |signed mode index|
            km = int(m[kept].max().item()) if bool(kept.any()) else -1
            if km < 0 or not torch.equal(kept, m <= km):
                raise NotImplementedError("de-aliasing mask is not a box in |k|; unsupported by the CUDA path")
            kmax.append(km)
        return kmax

    def wavenumber_tables(self):
        """1-D tables the kernels consume: dkraw_i = Im(2*pi*i*f_i), dk_i = same with the Nyquist
        entry zeroed (the Hermitian projection of a first-derivative symbol)."""
        dkraw, dk = [], []
        for i, n in enumerate(self.shape):
            r = (2j * torch.pi * self._f[i]).imag.contiguous()
            d = r.clone()
            if n % 2 == 0:
                d[n // 2] = 0
            dkraw.append(r)
            dk.append(d)
        return dk, dkraw
