"""Gradient mode: ``operator.integrate(u_0)`` / ``operator(u)`` with ``u.requires_grad`` for NONLINEAR operators.

The reference is differentiable because every step is a chain of torch ops (README.md:82; operator/_base.py:676-790 run
under autograd). The fused step kernels keep no graph, so when a gradient is asked for the step is unrolled instead:
every transform and every symbol product still runs on the library's kernels (``fsm_c2r``, ``fsm_r2c``,
``fsm_spectral_map``), each wrapped in a ``torch.autograd.Function`` whose backward is the adjoint pass on the same
kernels; the point-wise products of the nonlinear cores and the integrator's stage algebra (the same tables the fused
plan uses) are torch element-wise ops on the rot-half state, which autograd differentiates. Forward values agree with
the fused kernels to rounding (tests/test_autograd_nonlinear.py); gradients are pinned against the reference's own
autograd (tests/golden_grad).

Adjoints on the half spectrum (real inner product Re sum conj(a) b; w_k = 1 on the k = 0 and Nyquist lines of the last
axis, 2 elsewhere; n = number of grid points):
    c2r^T g = (w / n) * r2c(g)        r2c^T G = c2r(n * G / w)        map(S)^T = map(S^H)
"""
from typing import Optional

import torch

from . import _cabi
from .integrator import RK_TABLEAUS

_RK4 = ([[1 / 2, 1 / 2], [1 / 2, 0, 1 / 2], [1, 0, 0, 1]], [1 / 6, 1 / 3, 1 / 3, 1 / 6])     # integrator/_rk.py:142-155


def _adjoint_terms(terms):
    """(out, in, powers, inverse-Laplacian power, coef) of S -> terms of S^H: channels swapped, (i k)^p conjugated."""
    return [(ci, co, pw, q, coef * (-1.0) ** (sum(pw) % 2)) for (co, ci, pw, q, coef) in terms]


class _C2R(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_hat, ops):
        ctx.ops = ops
        return ops.plan(x_hat.shape[1]).c2r(x_hat.detach().contiguous())

    @staticmethod
    def backward(ctx, g):
        ops = ctx.ops
        return ops.plan(g.shape[1]).r2c(g.detach().contiguous()) * ops.w_over_n, None


class _R2C(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, ops):
        ctx.ops = ops
        return ops.plan(u.shape[1]).r2c(u.detach().contiguous())

    @staticmethod
    def backward(ctx, g_hat):
        ops = ctx.ops
        return ops.plan(g_hat.shape[1]).c2r((g_hat.detach() * ops.n_over_w).contiguous()), None


class _Map(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_hat, ops, c_out, terms, dealias):
        ctx.ops, ctx.c_in, ctx.terms, ctx.dealias = ops, x_hat.shape[1], terms, dealias
        return ops.plan(x_hat.shape[1]).spectral_map(x_hat.detach().contiguous(), c_out, terms, dealias)

    @staticmethod
    def backward(ctx, g_hat):
        ops = ctx.ops
        return ops.plan(g_hat.shape[1]).spectral_map(g_hat.detach().contiguous(), ctx.c_in, _adjoint_terms(ctx.terms),
                                                     ctx.dealias), None, None, None, None


class DiffOps:
    """Differentiable transforms and symbol products on the rot-half layout of one mesh / batch size."""

    def __init__(self, op, batch: int):
        self.op, self.B = op, batch
        f_mesh = op._state_dict["f_mesh"]
        self.shape = tuple(f_mesh.shape)
        self.d = len(self.shape)
        nh = self.shape[-1] // 2 + 1
        w = torch.full((nh,), 2.0, dtype=f_mesh.dtype, device=f_mesh.device)
        w[0] = 1.0
        if self.shape[-1] % 2 == 0:
            w[-1] = 1.0
        if self.d == 2:                                   # rot-half: [nh][n0]
            w = w[:, None].expand(nh, self.shape[0])
        elif self.d == 3:                                 # rot-half: [n1][nh][n0]
            w = w[None, :, None].expand(self.shape[1], nh, self.shape[0])
        n = 1
        for s in self.shape:
            n *= s
        w = w.reshape(-1)
        self.w_over_n, self.n_over_w = (w / n).contiguous(), (n / w).contiguous()
        self._e = [tuple(1 if i == a else 0 for i in range(3)) for a in range(3)]

    def plan(self, n_channel: int):
        return self.op._tf(self.B, n_channel)

    def c2r(self, x_hat):
        return _C2R.apply(x_hat, self)

    def r2c(self, u):
        return _R2C.apply(u, self)

    def smap(self, x_hat, c_out, terms, dealias=False):
        return _Map.apply(x_hat, self, c_out, terms, dealias)

    def mask(self, x_hat):
        C = x_hat.shape[1]
        return self.smap(x_hat, C, [(c, c, (0, 0, 0), 0, 1.0) for c in range(C)], True)

    def e(self, a):
        return self._e[a]

    # ---- layout converters as torch ops (what fsm_half_to_full / fsm_full_to_half compute) ---------------------
    def _unrot(self, x_hat):
        """(B, C, modes) rot-half -> (B, C, *shape[:-1], nh)"""
        shape, nh = self.shape, self.shape[-1] // 2 + 1
        lead = x_hat.shape[:-1]
        if self.d == 1:
            return x_hat.reshape(*lead, nh)
        if self.d == 2:
            return x_hat.reshape(*lead, nh, shape[0]).permute(0, 1, 3, 2)
        return x_hat.reshape(*lead, shape[1], nh, shape[0]).permute(0, 1, 4, 2, 3)

    def _mirror(self, t):
        """t(k) -> conj t(-k) over the grid axes"""
        dims = list(range(2, 2 + self.d))
        return torch.roll(torch.flip(t, dims), [1] * self.d, dims).conj()

    def half_to_full(self, x_hat):
        """Hermitian extension of a half spectrum: (B, C, modes) -> (B, C, N...)"""
        nh, nl = self.shape[-1] // 2 + 1, self.shape[-1]
        s = self._unrot(x_hat)
        rest = s[..., 1:nl - nh + 1].flip(-1).conj()
        if self.d > 1:
            dims = list(range(2, 1 + self.d))
            rest = torch.roll(torch.flip(rest, dims), [1] * len(dims), dims)
        return torch.cat([s, rest], dim=-1)

    def full_to_half(self, full_hat):
        """Hermitian projection of ANY full spectrum, stored half in the rot-half layout"""
        nh = self.shape[-1] // 2 + 1
        h = (0.5 * (full_hat + self._mirror(full_hat)))[..., :nh]
        if self.d == 2:
            h = h.permute(0, 1, 3, 2)
        elif self.d == 3:
            h = h.permute(0, 1, 3, 4, 2)
        return h.reshape(h.shape[0], h.shape[1], -1)


class _HostView:
    """What a host-composed core (``OperatorLike._external_nonlinear``) sees in place of the stepper: the same method
    names, differentiable."""

    def __init__(self, ops: DiffOps, n_channel: int):
        self.ops, self.B, self.C, self.local_shape = ops, ops.B, n_channel, ops.shape

    def c2r(self, x_hat):
        return self.ops.c2r(x_hat)

    def r2c(self, u):
        return self.ops.r2c(u)

    def spectral_map(self, x_hat, c_out, terms, dealias=False):
        return self.ops.smap(x_hat, c_out, terms, dealias)

    def mask_state(self, x_hat):
        return self.ops.mask(x_hat)

    def sym_outer(self, u):
        C = u.shape[1]
        return torch.stack([u[:, a] * u[:, c] for a in range(C) for c in range(a, C)], dim=1)

    def other(self, n_channel):
        return _HostView(self.ops, n_channel)

    # full spectra for user-defined NonlinearFunc cores, as torch ops so that autograd sees them
    def half_to_full(self, x_hat):
        return self.ops.half_to_full(x_hat)

    def full_to_half(self, full_hat):
        return self.ops.full_to_half(full_hat)

    @property
    def mesh(self):
        """What a user core receives as ``f_mesh`` in gradient mode: the same tables, differentiable transforms."""
        return _DiffMesh(self.ops)


class _DiffMesh:
    """``FourierMesh`` as seen by a user-defined core while a graph is being recorded: every table attribute is the
    mesh's own; ``fft`` / ``ifft`` (mesh.py:481-491) go through the differentiable passes of ``DiffOps``."""

    def __init__(self, ops: "DiffOps"):
        self._ops, self._mesh = ops, ops.op._state_dict["f_mesh"]

    def __getattr__(self, name):
        return getattr(self._mesh, name)

    def fft(self, u):
        if u.is_complex():
            return self.fft(u.real) + 1j * self.fft(u.imag)
        o, shape = self._ops, self._mesh.shape
        lead = u.shape[:u.dim() - len(shape)]
        flat = u.reshape(-1, 1, *shape)
        return _ForBatch(o, flat.shape[0]).fft(flat).reshape(*lead, *shape)

    def ifft(self, u_fft):
        o, shape = self._ops, self._mesh.shape
        lead = u_fft.shape[:u_fft.dim() - len(shape)]
        flat = u_fft.reshape(-1, 1, *shape)
        return _ForBatch(o, flat.shape[0]).ifft(flat).reshape(*lead, *shape)


class _ForBatch:
    """Differentiable full-spectrum transforms of ``n`` scalar fields (leading axes of a user core's argument flattened)."""

    def __init__(self, ops: "DiffOps", n: int):
        self.ops = ops if n == ops.B else DiffOps(ops.op, n)

    def fft(self, u):
        return self.ops.half_to_full(self.ops.r2c(u))

    def ifft(self, full):
        o = self.ops
        return torch.complex(o.c2r(o.full_to_half(full)), o.c2r(o.full_to_half(-1j * full)))


class GradientMode:
    """One operator on one mesh, batch size and time step, unrolled for autograd. ``st`` is the stepper the forward
    path would use: its tables (``rot_tables``: exp, half_exp, coef_i, lin in the rot-half layout) and constant
    source are reused as they are, so both modes integrate with identical coefficients."""

    def __init__(self, op, st):
        lo = op._lowered
        if st.P > 1:
            raise NotImplementedError("gradients are not available on slab-decomposed grids")
        if lo.get("force_hat") is not None or lo.get("dyn_force") is not None:
            raise NotImplementedError("NSPressureConvection with an external force is not differentiable on the CUDA path")
        if getattr(st, "ks_group", None) is not None:
            raise NotImplementedError("gradients are not available for ensembles sharded over a process group")
        self.op, self.st, self.lo = op, st, lo
        self.ops = DiffOps(op, st.B)
        self.C, self.d = st.C, self.ops.d
        self.name, self.dt = st.integrator, st.dt
        self.program = lo["program"]
        coef = lo.get("nl_coef_b")
        self.nl_coef = lo["nl_coef"] if coef is None else \
            coef.reshape(-1, 1, 1).to(device=st.device, dtype=st.rdtype) * lo["nl_coef"]
        self.external = op._external_nonlinear(lo["external"], st.C) if lo["external"] else None
        self.view = _HostView(self.ops, st.C)
        self.source = getattr(st, "source_rot", None)
        if self.source is not None:
            self.source = self.source.reshape(1, st.C, st.nmodes)

    def table(self, k) -> Optional[torch.Tensor]:
        t = self.st.rot_tables.get(k)
        return None if t is None else t.reshape(self.st.tab_batch, -1, self.st.nmodes)

    # ---- nonlinear cores on the rot-half state ----------------------------------------------------
    def _convection(self, x):
        """(u . grad) u on the dealiased state (generic/_convection.py:18-48)"""
        o, C, d = self.ops, self.C, self.d
        u = o.c2r(o.mask(x))
        # one map per transported channel: d gradient fields each (fsm_spectral_map takes at most 6 channels)
        adv = [(u * o.c2r(o.smap(x, d, [(j, c, o.e(j), 0, 1.0) for j in range(d)], True))).sum(dim=1) for c in range(C)]
        return o.r2c(torch.stack(adv, dim=1))

    def nonlinear(self, x):
        o, d, p = self.ops, self.d, self.program
        out = None
        if p == _cabi.PROG_CONVECTION:
            out = self._convection(x)
        elif p == _cabi.PROG_KS:                                  # dedicated/_ks_convection.py:18-38
            g = o.c2r(o.smap(x, d, [(j, 0, o.e(j), 0, 1.0) for j in range(d)], True))
            v = 0.5 * (g * g).sum(dim=1, keepdim=True)
            if self.lo["ks_remove_mean"]:
                v = v - v.mean()
            out = o.r2c(v)
        elif p == _cabi.PROG_NS2D_VORT:                                # dedicated/_navier_stokes.py:27-46
            f = o.c2r(o.smap(x, 4, [(0, 0, (0, 1, 0), 1, -1.0), (1, 0, (1, 0, 0), 1, 1.0),
                                    (2, 0, (1, 0, 0), 0, 1.0), (3, 0, (0, 1, 0), 0, 1.0)], True))
            out = o.r2c(f[:, 0:1] * f[:, 2:3] + f[:, 1:2] * f[:, 3:4])
        elif p == _cabi.PROG_NS3D:                                # dedicated/_navier_stokes.py:231-254 (no force)
            c = self._convection(x)
            terms = [(i, j, tuple(a + b for a, b in zip(o.e(i), o.e(j))), 1, 1.0) for i in range(d) for j in range(d)]
            terms += [(i, i, (0, 0, 0), 0, -1.0) for i in range(d)]
            out = o.smap(c, d, terms, False)
        if out is not None:
            out = out * self.nl_coef
        if self.external is not None:
            r = self.external(self.view, x)
            out = r if out is None else out + r
        if self.source is not None:
            out = self.source.expand(x.shape) if out is None else out + self.source
        if out is None:
            out = torch.zeros_like(x)
        return out

    def rhs(self, x):
        """L x + N(x)   (operator/_base.py:408-439)"""
        lin = self.table("lin")
        n = self.nonlinear(x)
        return n if lin is None else lin * x + n

    # ---- one step of the installed integrator ---------------------------------------------------------
    def step(self, x):
        T, N, name = self.table, self.nonlinear, self.name
        if name in RK_TABLEAUS or name == "RK4":                  # integrator/_rk.py:43-58
            rows, b = RK_TABLEAUS.get(name, _RK4)
            ks = [self.rhs(x)]
            for row in rows:
                ks.append(self.rhs(x + sum(self.dt * a * k for a, k in zip(row[1:], ks) if a != 0)))
            return x + sum(self.dt * bi * k for bi, k in zip(b, ks) if bi != 0)
        if name == "ETDRK0":                                       # integrator/_etdrk.py:10-82
            return T("exp") * x
        if name in ("ETDRK1", "SETDRK1"):                         # _setdrk_step.py:5-11
            return T("exp") * x + T("coef_1") * N(x)
        if name in ("ETDRK2", "SETDRK2"):                         # _setdrk_step.py:14-27
            n0 = N(x)
            a = T("exp") * x + T("coef_1") * n0
            return a + T("coef_2") * (N(a) - n0)
        if name == "SETDRK3":                                      # _setdrk_step.py:30-52
            n0 = N(x)
            n1 = N(T("half_exp") * x + T("coef_1") * n0)
            n2 = N(T("exp") * x + T("coef_2") * (2 * n1 - n0))
            return T("exp") * x + T("coef_3") * n0 + T("coef_4") * n1 + T("coef_5") * n2
        if name == "SETDRK4":                                      # _setdrk_step.py:55-82
            n0 = N(x)
            a = T("half_exp") * x + T("coef_1") * n0
            n1 = N(a)
            n2 = N(T("half_exp") * x + T("coef_2") * n1)
            n3 = N(T("half_exp") * a + T("coef_3") * (2 * n2 - n0))
            return T("exp") * x + T("coef_4") * n0 + T("coef_5") * (2 * (n1 + n2)) + T("coef_6") * n3
        raise NotImplementedError(name)

    def advance(self, x, n_steps: int):
        """``n_steps`` unrolled steps; with ``Operator.set_gradient_checkpointing(True)`` only the state at every step
        boundary is kept and each step is recomputed during backward (long rollouts on large grids)."""
        if getattr(self.op, "_grad_checkpoint", False) and torch.is_grad_enabled():
            from torch.utils.checkpoint import checkpoint
            for _ in range(int(n_steps)):
                x = checkpoint(self.step, x, use_reentrant=False)
            return x
        for _ in range(int(n_steps)):
            x = self.step(x)
        return x

    def integrate(self, u_0, n_steps: int):
        return self.ops.c2r(self.advance(self.ops.r2c(u_0.to(self.st.rdtype)), n_steps))

    def evaluate(self, u):
        return self.ops.c2r(self.rhs(self.ops.r2c(u.to(self.st.rdtype))))
