"""Spectral tables of the oracle (numpy). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Restates ``torchfsm/mesh.py`` of the reference:
  * per-axis frequencies            mesh.py:178-192  (``torch.fft.fftfreq(n, L/n)``)
  * broadcast views (1,1,..n_i..)   mesh.py:244-254
  * stacked vector (1,d,N...)       mesh.py:256-266
  * symbols (2*pi*i*f)^order        mesh.py:399-404
  * Laplacian / inverse Laplacian   mesh.py:406-426
  * nabla_vector                    mesh.py:436-441
  * 2/3-rule low-pass mask          mesh.py:443-461
  * fftn / ifftn over the mesh axes mesh.py:481-491
"""

import numpy as np
import scipy.fft as sfft

_REAL = {"float32": np.float32, "float64": np.float64}
_CPLX = {"float32": np.complex64, "float64": np.complex128}


class OracleMesh:
    """Fourier mesh over a periodic box; ``mesh_info`` = [(start, end, n), ...]."""

    def __init__(self, mesh_info, dtype="float32", workers=1):
        self.mesh_info = [tuple(m) for m in mesh_info]
        self.n_dim = len(self.mesh_info)
        self.dtype_name = np.dtype(dtype).name
        self.rdtype = _REAL[self.dtype_name]
        self.cdtype = _CPLX[self.dtype_name]
        self.shape = tuple(m[2] for m in self.mesh_info)
        self.fft_axes = tuple(range(-self.n_dim, 0))
        self.workers = workers
        self._f = [self._fftfreq(n, (b - a) / n) for (a, b, n) in self.mesh_info]

    def _fftfreq(self, n, d):
        # mesh.py:185 — torch.fft.fftfreq(n, d, dtype): integer ramp (0..ceil(n/2)-1,
        # -floor(n/2)..-1) scaled by 1/(n*d) in the working dtype (Nyquist negative).
        ramp = np.concatenate([np.arange(0, (n + 1) // 2), np.arange(-(n // 2), 0)])
        return (ramp.astype(self.rdtype) * self.rdtype(1.0 / (n * d))).astype(self.rdtype)

    # -- frequency views -------------------------------------------------
    def f(self, i):
        return self._f[i]

    def bf(self, i):
        shape = [1] * (self.n_dim + 2)
        shape[i + 2] = self.shape[i]
        return self._f[i].reshape(shape)

    def bf_vector(self):
        full = (1, 1) + self.shape
        return np.concatenate([np.broadcast_to(self.bf(i), full) for i in range(self.n_dim)], axis=1)

    # -- symbols -----------------------------------------------------------
    def grad(self, dim_i, order):
        # mesh.py:399-404: (2j*pi*bf)**order, complex working dtype.
        base = (self.cdtype(2j * np.pi) * self.bf(dim_i)).astype(self.cdtype)
        return _ipow(base, order)

    def nabla(self, order=1):
        # mesh.py:421-426: python sum() over axes => 0 + g_0 + g_1 + ...
        acc = 0
        for i in range(self.n_dim):
            acc = acc + self.grad(i, order)
        return acc

    def laplacian(self):
        return self.nabla(2)

    def invert_laplacian(self):
        # mesh.py:413-419: where(lap == 0, 1, 1/lap)
        lap = self.laplacian()
        safe = np.where(lap == 0, self.cdtype(1), lap)
        return np.where(lap == 0, self.cdtype(1), self.cdtype(1) / safe).astype(self.cdtype)

    def nabla_vector(self, order):
        base = (self.cdtype(2j * np.pi) * self.bf_vector()).astype(self.cdtype)
        return _ipow(base, order)

    def low_pass_filter(self, rel_freq_threshold=2 / 3):
        # mesh.py:443-461: product over axes of [|f_i| <= rate * max|f_i|].
        mask = np.ones((1, 1) + self.shape, dtype=self.rdtype)
        for i in range(self.n_dim):
            abs_f = np.abs(self.bf(i))
            # the comparison threshold is formed in the working dtype, as torch does
            thr = self.rdtype(abs_f.max()) * self.rdtype(rel_freq_threshold)
            mask = mask * np.where(abs_f > thr, self.rdtype(0), self.rdtype(1))
        return mask.astype(self.rdtype)

    # -- transforms ----------------------------------------------------------
    def fft(self, u):
        # mesh.py:481-485 — full C2C spectrum, norm="backward"
        return sfft.fftn(np.asarray(u).astype(self.cdtype, copy=False), axes=self.fft_axes,
                         workers=self.workers).astype(self.cdtype, copy=False)

    def ifft(self, u_hat):
        # mesh.py:487-491 — 1/N on the inverse
        return sfft.ifftn(np.asarray(u_hat).astype(self.cdtype, copy=False), axes=self.fft_axes,
                          workers=self.workers).astype(self.cdtype, copy=False)


def _ipow(z, order):
    """Integer power of a complex array the way torch evaluates ``z ** order``.

    torch special-cases exponent 2 as z*z (imaginary part of (i a)^2 is exactly 0,
    SURVEY.md §8 a4 [probe]); other exponents go through the generic complex pow.
    Order 1 is the identity. Orders >= 3 (KdV only, out of scope for the CUDA path)
    use numpy's pow and are not bit-pinned.
    """
    if order == 1:
        return z
    if order == 2:
        return z * z
    return z ** order
