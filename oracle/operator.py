"""Operator driver of the oracle (numpy). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Restates the reference's driver on the full complex spectrum:
  register_mesh: split generators into linear / nonlinear     operator/_base.py:581-624
  _build_linear_coefs: L = sum_i coef_i * core_i               operator/_base.py:339-357
  _build_nonlinear_funcs: dealias once, dispatch, accumulate   operator/_base.py:359-406
  _build_operator (RK right-hand side)                         operator/_base.py:408-439
  _build_integrator ("auto": linear -> ETDRK0 else SETDRK4)    operator/_base.py:441-526
  integrate: fft(u_0); step x forward; ifft(.).real            operator/_base.py:676-751
  __call__                                                     operator/_base.py:753-790
"""

import numpy as np

from .spectral import OracleMesh
from .cores import LINEAR_KINDS, NONLINEAR_KINDS, linear_core, nonlinear_core
from . import integrators as _int


class OracleOperator:
    """``terms`` = [(kind, coef, params_dict), ...]; see oracle/cores.py for the kinds."""

    def __init__(self, terms, de_aliasing_rate=2 / 3):
        self.terms = [(k, c, dict(p or {})) for (k, c, p) in terms]
        self.de_aliasing_rate = de_aliasing_rate
        self.mesh = None
        self.integrator = None
        self._integrator_name = "auto"
        self._integrator_cfg = {}

    # ------------------------------------------------------------------ registration
    def register_mesh(self, mesh_info, n_channel, dtype="float32", workers=1):
        mesh = mesh_info if isinstance(mesh_info, OracleMesh) else OracleMesh(mesh_info, dtype, workers)
        self.mesh, self.n_channel = mesh, n_channel
        lin, non = [], []
        for kind, coef, params in self.terms:
            if kind in LINEAR_KINDS:
                lin.append((coef, linear_core(kind, mesh, n_channel, params)))
            elif kind in NONLINEAR_KINDS:
                if isinstance(params.get("force"), OracleOperator):      # a force operator sees the same mesh and state
                    fop = params["force"]
                    fop.de_aliasing_rate = self.de_aliasing_rate
                    fop.register_mesh(mesh, n_channel)
                core = nonlinear_core(kind, mesh, n_channel, params)
                core.rate = self.de_aliasing_rate
                non.append((coef, core))
            else:
                raise ValueError(f"Operator {kind} is not supported")
        # _base.py:339-357 — python sum() => 0 + c_0*core_0 + c_1*core_1 ...
        self.linear_coef = None
        if lin:
            acc = 0
            for coef, core in lin:
                acc = acc + _scale(coef, core, mesh)
            self.linear_coef = acc.astype(mesh.cdtype)
        self.nonlinear_func = self._make_nonlinear(non) if non else None
        self.integrator = None
        return self

    def _make_nonlinear(self, non):
        mesh = self.mesh
        mask = mesh.low_pass_filter(self.de_aliasing_rate)

        def nonlinear_all(u_hat):                                       # _base.py:375-403
            result = 0.0
            d_hat = d_u = u = None
            for coef, fun in non:
                if fun.dealias:
                    if d_hat is None:
                        d_hat = u_hat * mask
                        d_u = mesh.ifft(d_hat).real
                    result = result + _scale(coef, fun(d_hat, mesh, d_u), mesh)
                else:
                    if u is None:
                        u = mesh.ifft(u_hat).real
                    result = result + _scale(coef, fun(u_hat, mesh, u), mesh)
            return result

        return nonlinear_all

    def rhs(self, u_hat):                                               # _base.py:408-439
        if self.nonlinear_func is None:
            return self.linear_coef * u_hat
        if self.linear_coef is None:
            return self.nonlinear_func(u_hat)
        return self.linear_coef * u_hat + self.nonlinear_func(u_hat)

    # -------------------------------------------------------------------- integrator
    def set_integrator(self, name, **cfg):
        self._integrator_name, self._integrator_cfg = name, cfg
        self.integrator = None

    def build_integrator(self, dt, tables=None):
        name = self._integrator_name
        if name == "auto":                                              # _base.py:451-455
            name = "ETDRK0" if self.nonlinear_func is None else "SETDRK4"
        if name == "RK4":
            self.integrator = _int.RK4(dt, self.rhs)
            return self.integrator
        if name in _int.RK_FAMILY:
            self.integrator = _int.explicit_rk(name, dt, self.rhs)
            return self.integrator
        cls = _int.INTEGRATORS[name]
        L = self.linear_coef
        if L is None:                                                   # _base.py:473-478
            L = np.zeros((1,), dtype=self.mesh.cdtype)
        if name == "ETDRK0":
            assert self.nonlinear_func is None, "The ETDRK0 integrator only supports linear term"
            self.integrator = cls(dt, L, tables=tables)
        elif name.startswith("ETDRK"):
            self.integrator = cls(dt, L, self.nonlinear_func, tables=tables)
        else:
            cfg = {k: v for k, v in self._integrator_cfg.items() if k != "cpu_cached"}
            self.integrator = cls(dt, L, self.nonlinear_func, tables=tables, **cfg)
        return self.integrator

    # ----------------------------------------------------------------------- driver
    def integrate(self, u_0=None, u_0_hat=None, dt=1.0, step=1, return_in_fourier=False, record_every=None):
        assert self.mesh is not None, "register_mesh first"
        if self.integrator is None or self.integrator.dt != dt:
            self.build_integrator(dt)
        u_hat = self.mesh.fft(u_0) if u_0_hat is None else np.asarray(u_0_hat).astype(self.mesh.cdtype)
        frames = []
        for i in range(step):                                           # _base.py:732-735
            if record_every and i % record_every == 0:
                frames.append(u_hat.copy())
            u_hat = self.integrator.step(u_hat)
        if record_every:
            frames.append(u_hat.copy())
            traj = np.stack(frames, axis=1)
            if return_in_fourier:
                return traj
            import scipy.fft as sfft
            axes = tuple(range(-self.mesh.n_dim, 0))
            return sfft.ifftn(traj, axes=axes).real.astype(self.mesh.rdtype)
        if return_in_fourier:
            return u_hat
        return self.mesh.ifft(u_hat).real.astype(self.mesh.rdtype)

    def __call__(self, u=None, u_hat=None, return_in_fourier=False):     # _base.py:753-790
        u_hat = self.mesh.fft(u) if u_hat is None else u_hat
        v = self.rhs(u_hat)
        return v if return_in_fourier else self.mesh.ifft(v).real.astype(self.mesh.rdtype)


def _scale(coef, x, mesh):
    """``coef * x`` with a python-float coefficient kept in the working precision (torch
    multiplies a tensor by a python scalar in the tensor's dtype)."""
    if isinstance(coef, (int, float)):
        return x * mesh.rdtype(coef)
    return np.asarray(coef) * x
