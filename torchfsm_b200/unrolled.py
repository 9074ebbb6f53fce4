"""Time stepping with a COMPLEX linear symbol (odd-order ``SpatialDerivative`` terms: advection, dispersion, the beta
effect) on 2-D and 3-D grids.

The reference carries a full complex spectrum (operator/_base.py:716-751). The symbol (2 pi i f)^n of an odd-order term
on its Nyquist plane (mesh.py:399-441; the Nyquist frequency is stored once, negative) is not conjugate-symmetric, so the
state leaves the Hermitian subspace there: S(k) and S(-k) evolve with tables T(k) and T(-k) that are not conjugates of
each other, and the final ``ifft(...).real`` returns the Hermitian part (S(k) + conj S(-k)) / 2. A half-spectrum state
cannot hold S(-k) on those planes, and the fused 2-D/3-D kernels take real tables only.

Exact restatement on half spectra: carry the PAIR X = (S, S~) with S~(k) = conj S(-k). S~ obeys the same integrator
with the mirrored tables T~(k) = conj T(-k) (equal to T away from the Nyquist planes). The nonlinear cores read the
dealiased state, whose Nyquist planes are masked (operator/_base.py:381-385), so both members see the same Hermitian
nonlinear term, evaluated once per stage on the library's kernels; the stage algebra runs on both members with their
own tables as torch element-wise ops (the machinery of gradient mode, autograd.py). Output = c2r((S + S~) / 2).
Requires a de-aliasing rate below 1 whenever a nonlinear term is present. Not fused: one launch per pass and per
stage operation; the fused kernels serve real symbols (and complex symbols on 1-D grids).
"""
from typing import Optional

import torch

from . import _cabi
from .autograd import GradientMode
from .integrator import build_tables, has_imag  # noqa: F401


def _mirror(t: torch.Tensor, n_dim: int) -> torch.Tensor:
    """t(k) -> conj t(-k) over the last ``n_dim`` axes of a full-layout tensor."""
    dims = list(range(t.dim() - n_dim, t.dim()))
    return torch.roll(torch.flip(t, dims), [1] * n_dim, dims).conj().resolve_conj()


def _rot(t: torch.Tensor, shape) -> torch.Tensor:
    """(..., *shape) full layout -> (..., rot-half modes): the layout of include/fsm_b200.h."""
    nd, nh = len(shape), shape[-1] // 2 + 1
    t = t[..., :nh]
    lead = list(range(t.dim() - nd))
    if nd == 2:
        t = t.permute(*lead, t.dim() - 1, t.dim() - 2)
    elif nd == 3:
        t = t.permute(*lead, t.dim() - 2, t.dim() - 1, t.dim() - 3)
    return t.reshape(*t.shape[:len(lead)], -1).contiguous()


def _unrot(t: torch.Tensor, shape) -> torch.Tensor:
    """(..., rot-half modes) -> (..., *shape[:-1], nh) full-layout half."""
    nd, nh = len(shape), shape[-1] // 2 + 1
    lead = t.shape[:-1]
    n = len(lead)
    if nd == 1:
        return t.reshape(*lead, nh)
    if nd == 2:
        return t.reshape(*lead, nh, shape[0]).permute(*range(n), n + 1, n)
    return t.reshape(*lead, shape[1], nh, shape[0]).permute(*range(n), n + 2, n, n + 1)


class _PairMode(GradientMode):
    def table(self, k):
        return self.st.pair_tables.get(k)

    def nonlinear(self, x):
        # both members share the dealiased (Hermitian) input of the nonlinear cores: one evaluation
        return super().nonlinear(x[0]).unsqueeze(0)


class PairedSpectrumStepper:
    """Integrator protocol (``.dt``, ``.step``, ``.forward``; operator/_base.py:462-491) and the native entry points of
    ``FusedStepper`` for operators whose linear symbol is complex on a 2-D/3-D grid. Every "half" state of this stepper is
    the pair (2, B, C, modes)."""

    P, slab, complex_tables, ks_group = 1, None, True, None

    def __init__(self, op, batch: int, name: str, dt: float, cfg: dict, tables: Optional[dict] = None):
        sd, lo = op._state_dict, op._lowered
        f_mesh, L = sd["f_mesh"], sd["linear_coef"]
        self.f_mesh, self.B, self.C, self.dt, self.integrator = f_mesh, batch, sd["n_channel"], dt, name
        self.shape = self.local_shape = tuple(f_mesh.shape)
        self.n_dim = len(self.shape)
        self.device, self.rdtype, self.cdtype = f_mesh.device, f_mesh.dtype, f_mesh.cdtype
        nonlinear = lo["program"] != _cabi.PROG_LINEAR or bool(lo["external"])
        if nonlinear and any(int(k) >= n // 2 for k, n in zip(lo["kmax"], self.shape)):
            raise NotImplementedError("a complex linear symbol (odd-order linear terms) with a nonlinear term needs a "
                                      "de-aliasing rate below 1 on 2-D/3-D grids: the nonlinear cores must not see the "
                                      "non-Hermitian content of the Nyquist planes")
        if any((t.kind == "implicit_func_source" and not t.params.get("non_linear", True)) or
               (t.kind == "custom_nonlinear" and not getattr(t.params["func"], "_dealiasing_swtich", True))
               for t in lo["external"]):
            raise NotImplementedError("a nonlinear core that reads the un-dealiased state (ImplicitSource(func, "
                                      "non_linear=False), NonlinearFunc(dealiasing_swtich=False)) is not available "
                                      "together with a complex linear symbol on 2-D/3-D grids")
        if getattr(op, "_ensemble_group", None) is not None or getattr(op, "_slab", None) is not None:
            raise NotImplementedError("complex linear symbols on 2-D/3-D grids run on one GPU only")
        self._tf = op._tf(batch, self.C)
        self.nmodes = self._tf.nmodes
        if tables is None:
            tables = build_tables(name, dt, L, **cfg)
        tabs = dict(tables)
        tabs["lin"] = L
        self.tables_full = tables
        self.pair_tables = {}
        for k, t in tabs.items():
            t = t.to(device=self.device)
            t = t.to(self.cdtype if t.is_complex() else self.rdtype)
            t = t.expand(t.shape[0], t.shape[1], *self.shape)
            if t.shape[0] not in (1, batch):
                raise ValueError("a batched coefficient must have one entry per sample")
            self.pair_tables[k] = torch.stack([_rot(t, self.shape), _rot(_mirror(t, self.n_dim), self.shape)])
        if lo["source_hat"] is not None:
            s = lo["source_hat"].to(device=self.device, dtype=self.cdtype)
            if s.shape[0] != 1:
                raise NotImplementedError("a per-sample explicit source is not supported by the CUDA path")
            self.source_rot = _rot(s[0].expand(self.C, *self.shape), self.shape)
        self._mode = _PairMode(op, self)

    # ---- native entry points on the pair ----------------------------------------------------------------------
    def empty_half(self) -> torch.Tensor:
        return torch.empty((2, self.B, self.C, self.nmodes), dtype=self.cdtype, device=self.device)

    def r2c(self, u: torch.Tensor) -> torch.Tensor:
        h = self._tf.r2c(u)
        return torch.stack([h, h])

    def c2r(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._tf.c2r(((x[0] + x[1]) * 0.5).contiguous(), out)

    def full_to_half(self, full_hat: torch.Tensor) -> torch.Tensor:
        """Any full spectrum, Hermitian or not (the reference's state between two steps), without loss: the stored half
        of S and of S~ (``fsm_full_to_half`` would project onto the Hermitian subspace)."""
        full_hat = full_hat.to(self.cdtype)
        return torch.stack([_rot(full_hat, self.shape), _rot(_mirror(full_hat, self.n_dim), self.shape)])

    def half_to_full(self, x: torch.Tensor) -> torch.Tensor:
        """The reference's (non-Hermitian) full spectrum: S(k) on the stored half; S(k) = conj S~(-k) on the other."""
        nh, nl = self.shape[-1] // 2 + 1, self.shape[-1]
        s, sm = _unrot(x[0], self.shape), _unrot(x[1], self.shape)
        full = torch.empty((self.B, self.C) + self.shape, dtype=self.cdtype, device=self.device)
        full[..., :nh] = s
        rest = sm[..., 1:nl - nh + 1].flip(-1).conj()                   # last-axis index j > n/2 <- n - j of S~
        if self.n_dim > 1:
            dims = list(range(2, 1 + self.n_dim))
            rest = torch.roll(torch.flip(rest, dims), [1] * len(dims), dims)
        full[..., nh:] = rest
        return full

    def step_half(self, x: torch.Tensor, n_steps: int = 1) -> torch.Tensor:
        with torch.no_grad():
            y = x
            for _ in range(int(n_steps)):
                y = self._mode.step(y)
            if y is not x:
                x.copy_(y)
        return x

    def rhs_half(self, x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self._mode.rhs(x).expand(x.shape)

    def step(self, u_hat_full: torch.Tensor) -> torch.Tensor:
        return self.half_to_full(self.step_half(self.full_to_half(u_hat_full), 1))

    def forward(self, u_hat_full: torch.Tensor, dt: float) -> torch.Tensor:
        return self.step(u_hat_full)

    # ---- gradient mode ----------------------------------------------------------------------------------------
    def integrate_with_grad(self, u_0: torch.Tensor, n_steps: int) -> torch.Tensor:
        ops = self._mode.ops
        h = ops.r2c(u_0.to(self.rdtype))
        x = self._mode.advance(torch.stack([h, h]), n_steps)
        return ops.c2r((x[0] + x[1]) * 0.5)

    def evaluate_with_grad(self, u: torch.Tensor) -> torch.Tensor:
        ops = self._mode.ops
        h = ops.r2c(u.to(self.rdtype))
        y = self._mode.rhs(torch.stack([h, h]))
        return ops.c2r((y[0] + y[1]) * 0.5)
