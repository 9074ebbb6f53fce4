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
Around the path (pinned by tests/golden_ops):
  grad (linear)                     operator/generic/_grad.py:6-15
  vorticity2velocity (linear)       operator/dedicated/_navier_stokes.py:75-92
  div, curl                         operator/generic/_div.py:9-23, _curl.py:9-55
  conservative_convection           operator/generic/_conservative_convection.py:8-27
  implicit_func_source              operator/generic/_source.py:20-42
  velocity2pressure / vorticity2pressure    operator/dedicated/_navier_stokes.py:119-163, 179-217
  ns_pressure_convection with external_force  :237-254, including the in-place `u_fft *= low_pass_filter()` on the
                                    caller's array and the force added twice
"""

import numpy as np

LINEAR_KINDS = ("laplacian", "biharmonic", "spatial_derivative", "implicit_unit_source", "grad", "vorticity2velocity")
NONLINEAR_KINDS = ("convection", "vorticity_convection", "ks_convection",
                   "ns_pressure_convection", "explicit_source", "div", "curl", "conservative_convection",
                   "implicit_func_source", "velocity2pressure", "vorticity2pressure")


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
    if kind == "grad":                                                   # _grad.py:12-15
        if n_channel != 1:
            raise ValueError("The Grad operator only supports scalar field.")
        return mesh.nabla_vector(1)
    if kind == "vorticity2velocity":                                     # _navier_stokes.py:81-92
        return _vorticity2velocity(mesh)
    raise ValueError(kind)


def _vorticity2velocity(mesh):
    full = (1, 1) + tuple(m[2] for m in mesh.mesh_info)
    return -1 * mesh.invert_laplacian() * np.concatenate(
        [np.broadcast_to(mesh.grad(1, 1), full), np.broadcast_to(-mesh.grad(0, 1), full)], axis=1)


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
    def __init__(self, force=None):
        self._conv = _Convection()
        self.force = force                       # an oracle operator (already registered) or None
        self.dealias = force is None             # _navier_stokes.py:227

    def __call__(self, u_hat, mesh, u=None):
        # _navier_stokes.py:237-254
        if self.force is not None:
            force = self.force.rhs(u_hat)        # evaluated on the un-dealiased state
            u_hat *= mesh.low_pass_filter(self.rate)     # IN PLACE on the caller's array (quirk Q5)
            u = mesh.ifft(u_hat).real
        elif u is None:
            u = mesh.ifft(u_hat).real
        conv = self._conv(u_hat, mesh, u)
        if self.force is not None:
            conv = conv - force
        nv = mesh.nabla_vector(1)
        p = mesh.invert_laplacian() * np.sum(nv * conv, axis=1, keepdims=True)
        if self.force is not None:
            return nv * p - conv + force         # the force enters twice, as in the reference
        return nv * p - conv


class _Div:
    dealias = False

    def __call__(self, u_hat, mesh, u=None):                             # _div.py:17-23
        return np.sum(mesh.nabla_vector(1) * u_hat, axis=1, keepdims=True)


class _Curl:
    dealias = False

    def __call__(self, u_hat, mesh, u=None):                             # _curl.py:17-55
        g = lambda i: mesh.grad(i, 1)          # noqa: E731
        if u_hat.shape[1] == 2:
            return g(0) * u_hat[:, 1:2] - g(1) * u_hat[:, 0:1]
        return np.concatenate([g(1) * u_hat[:, 2:3] - g(2) * u_hat[:, 1:2], g(2) * u_hat[:, 0:1] - g(0) * u_hat[:, 2:3],
                               g(0) * u_hat[:, 1:2] - g(1) * u_hat[:, 0:1]], axis=1)


class _ConservativeConvection:
    dealias = True

    def __call__(self, u_hat, mesh, u=None):                             # _conservative_convection.py:18-27
        if u is None:
            u = mesh.ifft(u_hat).real
        uu_hat = mesh.fft(u[:, :, None] * u[:, None])
        return (mesh.nabla_vector(1)[:, :, None] * uu_hat).sum(1)


class _ImplicitFuncSource:
    def __init__(self, func, non_linear=True):
        self.func, self.dealias = func, bool(non_linear)                 # _source.py:25-31

    def __call__(self, u_hat, mesh, u=None):                             # _source.py:33-42
        if u is None:
            u = mesh.ifft(u_hat).real
        return mesh.fft(np.asarray(self.func(u)))


class _ToPressure:
    """_Velocity2PressureCore / _Vorticity2PressureCore (_navier_stokes.py:119-163, 179-217)."""

    def __init__(self, from_vorticity, force=None):
        self._conv = _Convection()
        self.from_vorticity, self.force = from_vorticity, force
        self.dealias = force is None

    def __call__(self, u_hat, mesh, u=None):
        force = self.force.rhs(u_hat) if (self.force is not None and not self.from_vorticity) else None
        if self.from_vorticity:
            vel_hat = u_hat * _vorticity2velocity(mesh)
            if self.force is not None:
                vel_hat = vel_hat * mesh.low_pass_filter(self.rate)
            conv = self._conv(vel_hat, mesh, mesh.ifft(vel_hat).real)
            if self.force is not None:
                conv = conv - self.force.rhs(u_hat)
        else:
            if self.force is not None:
                u_hat *= mesh.low_pass_filter(self.rate)
                u = mesh.ifft(u_hat).real
            elif u is None:
                u = mesh.ifft(u_hat).real
            conv = self._conv(u_hat, mesh, u)
            if force is not None:
                conv = conv - force
        p = np.sum(mesh.nabla_vector(1) * conv, axis=1, keepdims=True)
        return -1 * p * mesh.invert_laplacian()


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
        return _NSPressureConvection(params.get("force"))
    if kind == "div":
        return _Div()
    if kind == "curl":
        return _Curl()
    if kind == "conservative_convection":
        return _ConservativeConvection()
    if kind == "implicit_func_source":
        return _ImplicitFuncSource(params["func"], params.get("non_linear", True))
    if kind in ("velocity2pressure", "vorticity2pressure"):
        return _ToPressure(kind == "vorticity2pressure", params.get("force"))
    if kind == "explicit_source":
        return _ExplicitSource(params["source"], mesh)
    raise ValueError(kind)
