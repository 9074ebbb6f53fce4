"""Reference-side adapter: run a GENUINE ``torchfsm`` operator (qiauil/torchfsm v0.0.4) on the CUDA library.

``lower(ref_operator, dt)`` inspects what the reference's own ``register_mesh`` produced (operator/_base.py:581-624)
— the linear coefficient tensor it built with its own expressions (``_state_dict["linear_coef"]``, :339-357) and the
nonlinear cores it collected (``_nonlinear_funcs``, pattern-matched by class: ``_ConvectionCore``,
``_ConservativeConvectionCore``, ``_VorticityConvectionCore``, ``_KSConvectionCore``, ``_NSPressureConvectionCore`` with its
external force, ``_ExplicitSourceCore``, ``_ImplicitFuncSourceCore``) —, translates them into an operator of this
package and lowers that with the package's own machinery. It returns an object that speaks the integrator protocol the hot loop expects (``.dt``, ``.step(u_hat)``,
``.forward(u_hat, dt)`` on full C2C spectra, operator/_base.py:462-491, 732-735) but advances the state with the
fused sm_100a kernels through the C ABI (include/fsm_b200.h).

``install(ref_operator)`` patches ``_build_integrator`` of that one operator instance so that the reference's
unmodified ``integrate`` loop uses it; anything the fused programs cannot express raises ``NotImplementedError``
inside ``lower`` and the patched method then falls back to the reference's own torch integrator — the fallback lives
on the REFERENCE side of the boundary, as INTEGRATION.md §2 describes; this package itself never falls back.

Nothing here imports ``torchfsm``: the adapter only looks at attributes of the objects it is handed.
"""

import torch

from .integrator import ETDRKIntegrator, SETDRKIntegrator, RKIntegrator
from .mesh import FourierMesh
from .operator import Operator, _Term

# reference core class -> term kind of this package (operator/generic/*.py, operator/dedicated/*.py, operator/_base.py)
_KIND_OF_CORE = {"_ConvectionCore": "convection", "_ConservativeConvectionCore": "conservative_convection",
                 "_KSConvectionCore": "ks_convection", "_VorticityConvectionCore": "vorticity_convection",
                 "_NSPressureConvectionCore": "ns_pressure_convection", "_ExplicitSourceCore": "explicit_source",
                 "_ImplicitFuncSourceCore": "implicit_func_source"}


def translate(ref_op, f_mesh: FourierMesh, n_channel: int) -> Operator:
    """A registered reference operator as an operator of this package: its linear coefficient tensor as it stands
    (``_state_dict["linear_coef"]``, built by the reference's own expressions, operator/_base.py:339-357) plus one term per
    nonlinear core it collected (``_nonlinear_funcs``, :359-406), recognised by class. An external force handed to
    ``NSPressureConvection`` is itself a reference operator and is translated the same way."""
    terms = []
    L = ref_op._state_dict.get("linear_coef")
    if L is not None:
        terms.append(_Term("linear_tensor", 1, {"L": L}))
    for coef, core in getattr(ref_op, "_nonlinear_funcs", []) or []:
        cls = type(core).__name__
        kind = _KIND_OF_CORE.get(cls)
        if kind is None:
            # a user-defined NonlinearFunc (operator/_base.py:56-103): called as it is, on full spectra the library's passes
            # produce, with this package's FourierMesh (same table and fft/ifft names); `lower` dry-runs it once
            if not callable(core) or not hasattr(core, "_dealiasing_swtich"):
                raise NotImplementedError(f"nonlinear core {cls} has no counterpart on the fused CUDA path")
            terms.append(_Term("custom_nonlinear", coef, {"func": core}))
            continue
        params = {}
        if kind == "ks_convection":
            params["remove_mean"] = bool(core.remove_mean)
        elif kind == "explicit_source":
            params.update(source=core.source, in_fourier=True)          # already fftn(source), operator/_base.py:1002-1005
        elif kind == "implicit_func_source":
            params.update(source_func=core.source_func, non_linear=bool(getattr(core, "_dealiasing_swtich", True)))
        elif kind == "ns_pressure_convection":
            force = getattr(core, "external_force", None)
            if force is not None:
                if force._state_dict.get("f_mesh") is None:             # the reference registers it at its first evaluation
                    force.register_mesh(ref_op._state_dict["f_mesh"], n_channel)
                params["external_force"] = translate(force, f_mesh, n_channel)
        elif not getattr(core, "_dealiasing_swtich", True):
            raise NotImplementedError("convective cores without de-aliasing are not supported")
        terms.append(_Term(kind, coef, params))
    if not terms:
        raise NotImplementedError("empty operator")
    return Operator(terms)


class LoweredIntegrator:
    """What ``install`` puts into ``_state_dict['integrator']`` of a reference operator. The reference's integrators
    are batch-agnostic; a plan is bound to a batch size, so steppers are created on first use per batch size."""

    def __init__(self, op: Operator, dt: float):
        self.dt, self._op, self._steppers = dt, op, {}

    def stepper(self, batch: int):
        st = self._steppers.get(batch)
        if st is None:
            st = self._steppers[batch] = self._op._build_integrator(self.dt, batch)
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
    """Translate a registered reference operator (``translate``) and lower it with this package's own machinery: fused
    programs, host-composed cores, external forces, per-sample coefficients, every ETD / RK integrator it supports. Raises
    ``NotImplementedError`` for anything outside (user ``NonlinearFunc`` subclasses, adaptive RK, unsupported grids, ...)."""
    sd = ref_op._state_dict
    ref_mesh = sd.get("f_mesh")
    if ref_mesh is None:
        raise ValueError("register_mesh must run before lowering (operator/_base.py:581-624)")
    n_channel = sd["n_channel"]
    f_mesh = FourierMesh([tuple(m) for m in ref_mesh.mesh_info], device=ref_mesh.device, dtype=ref_mesh.dtype)
    cfg = dict(getattr(ref_op, "_integrator_config", {}) or {})
    if cfg.get("adaptive"):
        raise NotImplementedError("adaptive Runge-Kutta stepping is host-synchronous and stays on the torch path")
    op = translate(ref_op, f_mesh, n_channel)
    op._de_aliasing_rate = getattr(ref_op, "_de_aliasing_rate", 2 / 3)
    solver = ref_op._integrator
    if isinstance(solver, str):
        op.set_integrator("auto")
    else:
        name = getattr(solver, "name", None)
        for enum in (ETDRKIntegrator, SETDRKIntegrator, RKIntegrator):
            if name in enum.__members__:
                op.set_integrator(enum[name], **{k: v for k, v in cfg.items()
                                                 if k in ("n_integration_points", "integration_radius", "cpu_cached")})
                break
        else:
            raise NotImplementedError(f"integrator {solver!r} is not supported by the fused CUDA path")
    op.register_mesh(f_mesh, n_channel)
    low = LoweredIntegrator(op, dt)
    # building a plan validates the configuration now (NotImplementedError if unsupported); per-sample coefficients fix
    # the batch, otherwise one sample is enough for the check
    L = sd.get("linear_coef")
    nb = L.shape[0] if (L is not None and L.shape[0] > 1) else 1
    low._steppers[nb] = op._build_integrator(dt, nb)
    if any(t.kind == "custom_nonlinear" for t in op.terms):
        st = low._steppers[nb]
        try:                                       # a user core may lean on reference internals this package does not have
            st.rhs_half(st.empty_half().zero_())
        except NotImplementedError:
            raise
        except Exception as e:                     # noqa: BLE001
            raise NotImplementedError(f"user-defined core failed on the CUDA path: {e!r}")
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
