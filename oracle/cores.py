"""Operator cores of the oracle (numpy). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

An operator is a list of *terms* ``(kind, coef, params)``. Linear kinds return a
coefficient tensor; nonlinear kinds are callables ``(u_hat, mesh, u) -> N_hat`` on
the reference's full complex spectrum ``(B, C, N...)``.

Linear cores restated:
  laplacian             operator/generic/_laplacian.py:7-15
  biharmonic            operator/generic/_biharmonic.py:8-16
  spatial_derivative    operator/generic/_spatial_derivative.py:7-20
  implicit_unit_source  operator/generic/_source.py:9-17   (ones_like(bf_x): shape (1,1,Nx,1,..))
Nonlinear cores restated:
  convection              operator/generic/_convection.py:18-48
  vorticity_convection    operator/dedicated/_navier_stokes.py:27-46
  ks_convection           operator/dedicated/_ks_convection.py:18-38
  ns_pressure_convection  operator/dedicated/_navier_stokes.py:231-254 (external_force=None)
  explicit_source         operator/_base.py:994-1015
"""

import numpy as np

LINEAR_KINDS = ("laplacian", "biharmonic", "spatial_derivative", "implicit_unit_source")
NONLINEAR_KINDS = ("convection", "vorticity_convection", "ks_convection",
                   "ns_pressure_convection", "explicit_source")


# ----------------------------------------------------------------------------- linear
def linear_core(kind, mesh, n_channel, params):
    if kind == "laplacian":
        return np.concatenate([mesh.laplacian()] * n_channel, axis=1)
    if kind == "biharmonic":
        lap = mesh.laplacian()
        return np.concatenate([lap * lap] * n_channel, axis=1)
    if kind == "spatial_derivative":
        return mesh.grad(params["dim_index"], params["order"])
    if kind == "implicit_unit_source":
        return np.ones_like(mesh.bf(0))
    raise ValueError(kind)


# --------------------------------------------------------------------------- nonlinear
class _Convection:
    dealias = True

    def __call__(self, u_hat, mesh, u=None):
        return mesh.fft(self.spatial_value(u_hat, mesh, u))

    def spatial_value(self, u_hat, mesh, u=None):
        # _convection.py:43-46: u . grad(u) with grad via ifft(i k_j u_hat_c).real
        if u is None:
            u = mesh.ifft(u_hat).real
        nabla_u = mesh.nabla_vector(1)[:, :, None] * u_hat[:, None]      # (B, d, C, N...)
        return (u[:, :, None] * mesh.ifft(nabla_u).real).sum(1)


class _VorticityConvection:
    dealias = True

    def __call__(self, w_hat, mesh, u=None):
        return mesh.fft(self.spatial_value(w_hat, mesh, u))

    def spatial_value(self, w_hat, mesh, u=None):
        # _navier_stokes.py:41-46
        psi = -w_hat * mesh.invert_laplacian()
        ux = mesh.ifft(mesh.grad(1, 1) * psi).real
        uy = mesh.ifft(-mesh.grad(0, 1) * psi).real
        gx = mesh.ifft(mesh.grad(0, 1) * w_hat).real
        gy = mesh.ifft(mesh.grad(1, 1) * w_hat).real
        return ux * gx + uy * gy


class _KSConvection:
    dealias = True

    def __init__(self, remove_mean=True):
        self.remove_mean = remove_mean

    def __call__(self, u_hat, mesh, u=None):
        return mesh.fft(self.spatial_value(u_hat, mesh, u))

    def spatial_value(self, u_hat, mesh, u=None):
        # _ks_convection.py:33-38; the mean spans batch AND space
        grad_u = mesh.ifft(mesh.nabla_vector(1) * u_hat).real
        re = mesh.rdtype(0.5) * np.sum(grad_u ** 2, axis=1, keepdims=True)
        if self.remove_mean:
            return re - re.mean(dtype=mesh.rdtype)
        return re


class _NSPressureConvection:
    dealias = True  # external_force is None => dealiased input (_navier_stokes.py:227)

    def __init__(self):
        self._conv = _Convection()

    def __call__(self, u_hat, mesh, u=None):
        # _navier_stokes.py:241-254 without external force
        if u is None:
            u = mesh.ifft(u_hat).real
        conv = self._conv(u_hat, mesh, u)
        nv = mesh.nabla_vector(1)
        p = mesh.invert_laplacian() * np.sum(nv * conv, axis=1, keepdims=True)
        return nv * p - conv


class _ExplicitSource:
    dealias = False

    def __init__(self, source, mesh):
        # _base.py:1002-1005: fftn of the spatial source over all trailing axes
        src = np.asarray(source)
        axes = tuple(range(2, src.ndim))
        import scipy.fft as sfft
        self.source_hat = sfft.fftn(src, axes=axes).astype(mesh.cdtype)

    def __call__(self, u_hat, mesh, u=None):
        return self.source_hat


def nonlinear_core(kind, mesh, n_channel, params):
    if kind == "convection":
        if mesh.n_dim != n_channel:
            raise ValueError("convection needs n_channel == n_dim")       # _convection.py:60-63
        return _Convection()
    if kind == "vorticity_convection":
        if mesh.n_dim != 2 or n_channel != 1:
            raise ValueError("Only vorticity in 2Dmesh is supported")     # _navier_stokes.py:57-58
        return _VorticityConvection()
    if kind == "ks_convection":
        return _KSConvection(params.get("remove_mean", True))
    if kind == "ns_pressure_convection":
        return _NSPressureConvection()
    if kind == "explicit_source":
        return _ExplicitSource(params["source"], mesh)
    raise ValueError(kind)
