"""Reference-side adapter: run a GENUINE ``torchfsm`` operator (qiauil/torchfsm v0.0.4) on the CUDA library.

``lower(ref_operator, dt)`` inspects what the reference's own ``register_mesh`` produced (operator/_base.py:581-624)
— the linear coefficient tensor it built with its own expressions (``_state_dict["linear_coef"]``, :339-357) and the
nonlinear cores it collected (``_nonlinear_funcs``, pattern-matched by class: ``_ConvectionCore``,
``_VorticityConvectionCore``, ``_KSConvectionCore``, ``_NSPressureConvectionCore``, ``_ExplicitSourceCore``) — and
returns an object that speaks the integrator protocol the hot loop expects (``.dt``, ``.step(u_hat)``,
``.forward(u_hat, dt)`` on full C2C spectra, operator/_base.py:462-491, 732-735) but advances the state with the
fused sm_100a kernels through the C ABI (include/fsm_b200.h).

``install(ref_operator)`` patches ``_build_integrator`` of that one operator instance so that the reference's
unmodified ``integrate`` loop uses it; anything the fused programs cannot express raises ``NotImplementedError``
inside ``lower`` and the patched method then falls back to the reference's own torch integrator — the fallback lives
on the REFERENCE side of the boundary, as INTEGRATION.md §2 describes; this package itself never falls back.

Nothing here imports ``torchfsm``: the adapter only looks at attributes of the objects it is handed.
"""

import torch

from . import _cabi
from .integrator import build_tables
from .mesh import FourierMesh
from .operator import FusedStepper

_PROGRAM_OF_CORE = {"_ConvectionCore": _cabi.PROG_CONVECTION, "_KSConvectionCore": _cabi.PROG_KS,
                    "_VorticityConvectionCore": _cabi.PROG_NS2D_VORT, "_NSPressureConvectionCore": _cabi.PROG_NS3D}


def _integrator_name(ref_op, linear: bool) -> str:
    solver = ref_op._integrator
    if isinstance(solver, str):                      # "auto": operator/_base.py:451-455
        return "ETDRK0" if linear else "SETDRK4"
    name = getattr(solver, "name", None)
    if name not in _cabi.INTEGRATOR_IDS:
        raise NotImplementedError(f"integrator {solver!r} is not supported by the fused CUDA path")
    return name


class LoweredIntegrator:
    """What ``install`` puts into ``_state_dict['integrator']`` of a reference operator. The reference's integrators
    are batch-agnostic; a plan is bound to a batch size, so steppers are created on first use per batch size."""

    def __init__(self, f_mesh: FourierMesh, n_channel: int, program: int, integrator: str, dt: float, linear_coef,
                 nl_coef: float, source_hat, kmax, ks_remove_mean: bool, cfg: dict):
        self.dt = dt
        self._args = (f_mesh, n_channel, program, integrator, dt, linear_coef, nl_coef, source_hat, kmax, ks_remove_mean, cfg)
        self._steppers = {}
        self._tables = None

    def stepper(self, batch: int, allocate: bool = True) -> FusedStepper:
        st = self._steppers.get(batch) if allocate else None
        if st is None:
            f_mesh, c, program, integ, dt, L, nl_coef, src, kmax, ks_mean, cfg = self._args
            if self._tables is None:                 # built once with the reference's expressions
                Lt = L
                if Lt is None:                        # operator/_base.py:473-478
                    Lt = torch.tensor([0.0], dtype=f_mesh.cdtype, device=f_mesh.device).reshape([1] * (f_mesh.n_dim + 2))
                self._tables = build_tables(integ, dt, Lt, **cfg)
            st = FusedStepper(f_mesh, batch, c, program, integ, dt, L, nl_coef, src, kmax, ks_mean, cfg,
                              tables=self._tables, allocate=allocate)
            if allocate:
                self._steppers[batch] = st
        return st

    def step(self, u_hat_full: torch.Tensor) -> torch.Tensor:
        return self.stepper(u_hat_full.shape[0]).step(u_hat_full)

    def forward(self, u_hat_full: torch.Tensor, dt: float) -> torch.Tensor:
        return self.step(u_hat_full)

    def steps(self, u_hat_full: torch.Tensor, n: int) -> torch.Tensor:
        """``n`` steps with the state kept in the half-spectrum layout in between (what a reference-side ``integrate``
        would call when no recorder is attached)."""
        st = self.stepper(u_hat_full.shape[0])
        return st.half_to_full(st.step_half(st.full_to_half(u_hat_full), n))


def lower(ref_op, dt: float) -> LoweredIntegrator:
    """Pattern-match a registered reference operator into a fused step program. Raises ``NotImplementedError`` for
    anything outside the fused programs (user ``NonlinearFunc`` subclasses, ``ImplicitSource(func)``, tensor
    coefficients on nonlinear terms, adaptive RK, ...)."""
    sd = ref_op._state_dict
    ref_mesh = sd.get("f_mesh")
    if ref_mesh is None:
        raise ValueError("register_mesh must run before lowering (operator/_base.py:581-624)")
    n_channel = sd["n_channel"]
    f_mesh = FourierMesh([tuple(m) for m in ref_mesh.mesh_info], device=ref_mesh.device, dtype=ref_mesh.dtype)
    if getattr(ref_op, "_integrator_config", {}).get("adaptive"):
        raise NotImplementedError("adaptive Runge-Kutta stepping is host-synchronous and stays on the torch path")
    program, nl_coef, ks_remove_mean, source_hat = _cabi.PROG_LINEAR, 0.0, True, None
    for coef, core in getattr(ref_op, "_nonlinear_funcs", []):
        cls = type(core).__name__
        if cls in _PROGRAM_OF_CORE:
            if program != _cabi.PROG_LINEAR:
                raise NotImplementedError("only one convective nonlinear term per operator is supported")
            if isinstance(coef, torch.Tensor):
                raise NotImplementedError("tensor-valued coefficients on nonlinear terms are not supported")
            program, nl_coef = _PROGRAM_OF_CORE[cls], float(coef)
            if cls == "_KSConvectionCore":
                ks_remove_mean = bool(core.remove_mean)
            if cls == "_NSPressureConvectionCore" and core.external_force is not None:
                raise NotImplementedError("NSPressureConvection with an external force is not supported")
            if not getattr(core, "_dealiasing_swtich", True):
                raise NotImplementedError("convective cores without de-aliasing are not supported")
        elif cls == "_ExplicitSourceCore":
            s_hat = coef * core.source.to(f_mesh.device)             # operator/_base.py:1002-1015
            source_hat = s_hat if source_hat is None else source_hat + s_hat
        else:
            raise NotImplementedError(f"nonlinear core {cls} has no fused program")
    L = sd.get("linear_coef")
    # an explicit source is a nonlinear core in the reference: "auto" picks SETDRK4 for it (operator/_base.py:451-455)
    name = _integrator_name(ref_op, program == _cabi.PROG_LINEAR and source_hat is None)
    if name == "ETDRK0" and program != _cabi.PROG_LINEAR:
        raise AssertionError("The ETDRK0 integrator only supports linear term")
    rate = getattr(ref_op, "_de_aliasing_rate", 2 / 3)
    kmax = f_mesh.low_pass_kmax(rate) if program != _cabi.PROG_LINEAR else [n // 2 for n in f_mesh.shape]
    if program == _cabi.PROG_NS3D and any(k >= n // 2 and n % 2 == 0 for k, n in zip(kmax, f_mesh.shape)):
        raise NotImplementedError("NSPressureConvection needs a de-aliasing rate below 1 on the fused path")
    cfg = {k: v for k, v in getattr(ref_op, "_integrator_config", {}).items()
           if k in ("n_integration_points", "integration_radius", "cpu_cached")}
    low = LoweredIntegrator(f_mesh, n_channel, program, name, dt, L, nl_coef, source_hat, kmax, ks_remove_mean, cfg)
    # plan creation validates the configuration now (NotImplementedError if unsupported); no workspace is allocated
    low.stepper(L.shape[0] if (L is not None and L.shape[0] > 1) else 1, allocate=False)   # per-sample coefficients fix the batch
    return low


def install(ref_op, strict: bool = False):
    """Patch ONE reference operator instance: its ``_build_integrator`` (operator/_base.py:441-526) installs a
    ``LoweredIntegrator`` when the registered cores match a fused program, and calls the reference's original
    method otherwise (``strict=True`` re-raises instead). Returns the operator."""
    original = ref_op._build_integrator

    def _build_integrator(dt):
        try:
            low = lower(ref_op, dt)
        except NotImplementedError:
            if strict:
                raise
            return original(dt)
        ref_op._state_dict["integrator"] = low
        ref_op._is_etdrk_integrator = True          # `integrate` then rebuilds only when dt changes (:721-725)
        return None

    ref_op._build_integrator = _build_integrator
    ref_op._b200_lowered = True
    return ref_op
