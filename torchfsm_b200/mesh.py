"""Mesh and spectral tables (host-side mirror of the reference's ``torchfsm/mesh.py``).

Everything here is set-up work done with torch ops on the target device; the hot path only
consumes the small 1-D tables and the coefficient tables derived from them. The formulas
follow the reference exactly (same expressions, same dtype, same order of operations) because
the ETD coefficient tables built from them are inputs to the CUDA path:
  per-axis frequencies  mesh.py:178-192   symbols (2*pi*i*f)^n  mesh.py:399-404
  Laplacian             mesh.py:406-426   2/3-rule mask         mesh.py:443-461
"""
from typing import Sequence, Union

import torch


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

    @property
    def cdtype(self):
        return torch.complex64 if self.dtype == torch.float32 else torch.complex128

    def f(self, i):
        return self._f[i]

    def bf(self, i):
        shape = [1] * (self.n_dim + 2)
        shape[i + 2] = self.shape[i]
        return self._f[i].reshape(shape)

    def grad(self, dim_i: int, order: int):
        return (2j * torch.pi * self.bf(dim_i)) ** order

    def nabla(self, order: int = 1):
        return sum(self.grad(i, order) for i in range(self.n_dim))

    def laplacian(self):
        return self.nabla(2)

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
