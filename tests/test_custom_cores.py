"""User-defined cores (the reference's extension protocol, operator/_base.py:16-131): ``LinearCoef``, ``NonlinearFunc``,
``CoreGenerator`` subclasses handed to ``Operator`` / ``LinearOperator`` / ``NonlinearOperator`` run on the library's passes
(emulator build here): against the same physics written with built-in operators, and against the unmodified reference
running the SAME core source where it is importable."""
import os
import sys

import pytest
import torch

from product_util import build_emulator


@pytest.fixture(scope="module", autouse=True)
def emulator():
    from torchfsm_b200 import _cabi
    prev = _cabi._lib
    _cabi.use_library(build_emulator())
    yield
    _cabi._lib = prev


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def make_cores(base):
    """The same user code against either package: ``base`` provides LinearCoef / NonlinearFunc / CoreGenerator."""

    class Cubic(base.NonlinearFunc):                      # -u^3 on the dealiased field
        def __call__(self, u_fft, f_mesh, u=None):
            if u is None:
                u = f_mesh.ifft(u_fft).real
            return f_mesh.fft(-u ** 3)

    class GradientSquared(base.NonlinearFunc):            # |grad u|^2 from the mesh's own symbols, un-dealiased input
        def __init__(self):
            super().__init__(dealiasing_swtich=False)

        def __call__(self, u_fft, f_mesh, u=None):
            g = f_mesh.ifft(f_mesh.nabla_vector(1) * u_fft).real          # (B, d, N...) from a (B, 1, N...) field
            return f_mesh.fft((g * g).sum(dim=1, keepdim=True))

    class HyperViscosity(base.LinearCoef):                # -nu lap^2, every channel
        def __init__(self, nu):
            self.nu = nu

        def __call__(self, f_mesh, n_channel):
            return torch.cat([-self.nu * f_mesh.laplacian() ** 2] * n_channel, dim=1)

    class ByChannels(base.CoreGenerator):                 # decides once the channel count is known
        def __call__(self, f_mesh, n_channel):
            if n_channel != 1:
                raise ValueError("scalar fields only")
            return Cubic()

    return Cubic, GradientSquared, HyperViscosity, ByChannels


def _u0(*shape):
    import torchfsm_b200 as fsm
    g = torch.Generator().manual_seed(17)
    u = torch.randn(*shape, dtype=torch.float64, generator=g)
    mesh = fsm.MeshGrid([(0, 1, s) for s in shape[2:]], dtype=torch.float64)
    u = (0.002 * fsm.Laplacian()).integrate(u, mesh=mesh, dt=1.0, step=1)
    return u / u.abs().max()


def test_custom_cores_equal_the_builtin_formulation():
    import torchfsm_b200 as fsm
    Cubic, GradientSquared, HyperViscosity, ByChannels = make_cores(fsm)
    mesh = fsm.MeshGrid([(0, 1, 16), (0, 1, 32)], dtype=torch.float64)
    u0 = _u0(2, 1, 16, 32)
    builtin = 0.05 * fsm.Laplacian() + fsm.ImplicitSource() + fsm.ImplicitSource(lambda u: -u ** 3)
    for custom in (0.05 * fsm.Laplacian() + fsm.ImplicitSource() + fsm.NonlinearOperator(Cubic()),
                   0.05 * fsm.Laplacian() + fsm.ImplicitSource() + fsm.Operator(ByChannels()),
                   fsm.Operator([fsm.Laplacian().terms[0], Cubic()], coefs=[0.05, 1]) + fsm.ImplicitSource()):
        for op in (builtin, custom):
            op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
        want = builtin.integrate(u0, mesh=mesh, dt=0.01, step=4)
        got = custom.integrate(u0, mesh=mesh, dt=0.01, step=4)
        assert type(custom._state_dict["integrator"]).__name__ == "HostComposedStepper"
        assert _rel(got, want) < 1e-13
        assert _rel(custom(u0, mesh=mesh), builtin(u0, mesh=mesh)) < 1e-13
    # a user-defined linear symbol: same tables as the built-in biharmonic operator, fused linear step
    lin_custom, lin_builtin = fsm.LinearOperator(HyperViscosity(1e-4)), -1e-4 * fsm.Biharmonic()
    assert lin_custom.is_linear is False                      # unknown before a mesh is registered, as in the reference
    got, want = (op.integrate(u0, mesh=mesh, dt=0.1, step=3) for op in (lin_custom, lin_builtin))
    assert lin_custom.is_linear and _rel(got, want) < 1e-14
    fresh = fsm.LinearOperator(HyperViscosity(1e-4))          # solve() registers the mesh itself
    assert _rel(fresh.solve(b=u0, mesh=mesh, n_channel=1), lin_builtin.solve(b=u0, mesh=mesh, n_channel=1)) < 1e-14
    with pytest.raises(ValueError):
        fsm.Operator(ByChannels()).integrate(_u0(1, 2, 16, 32), mesh=mesh, dt=0.01, step=1)
    with pytest.raises(ValueError):
        fsm.Operator([Cubic()], coefs=[1, 2])


def test_custom_cores_match_the_reference_running_the_same_source():
    if not os.path.isdir("/root/reference/torchfsm"):
        pytest.skip("reference not present")
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import torchfsm.operator as rop
    import torchfsm.operator._base as rbase
    import torchfsm.integrator as rint
    from torchfsm.mesh import MeshGrid as RefMesh
    import torchfsm_b200 as fsm
    mi = [(0, 1, 16), (0, 1, 32)]
    u0 = _u0(2, 1, 16, 32)
    outs = []
    for pkg, base, mesh, etd in ((rop, rbase, RefMesh(mi, dtype=torch.float64), rint.SETDRKIntegrator),
                                 (fsm, fsm, fsm.MeshGrid(mi, dtype=torch.float64), fsm.SETDRKIntegrator)):
        Cubic, GradientSquared, HyperViscosity, ByChannels = make_cores(base)
        op = 0.05 * pkg.Laplacian() + pkg.LinearOperator(HyperViscosity(1e-5)) + pkg.NonlinearOperator(Cubic()) \
            + 0.3 * pkg.NonlinearOperator(GradientSquared())
        op.set_integrator(etd.SETDRK3)
        outs.append((op.integrate(u0.clone(), mesh=mesh, dt=0.005, step=3), op(u0.clone(), mesh=mesh)))
    assert _rel(outs[1][0], outs[0][0]) < 1e-12 and _rel(outs[1][1], outs[0][1]) < 1e-12


def test_reference_operator_with_a_user_core_through_the_adapter():
    if not os.path.isdir("/root/reference/torchfsm"):
        pytest.skip("reference not present")
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import copy
    import torchfsm.operator as rop
    import torchfsm.operator._base as rbase
    from torchfsm.mesh import MeshGrid as RefMesh
    from torchfsm_b200 import reference_adapter
    Cubic, GradientSquared, _, _ = make_cores(rbase)
    mesh = RefMesh([(0, 1, 16), (0, 1, 32)], dtype=torch.float64)
    u0 = _u0(2, 1, 16, 32)
    op = 0.05 * rop.Laplacian() + rop.NonlinearOperator(Cubic()) + 0.3 * rop.NonlinearOperator(GradientSquared())
    want = copy.deepcopy(op).integrate(u0, mesh=mesh, dt=0.005, step=3)
    fused = reference_adapter.install(copy.deepcopy(op), strict=True)
    got = fused.integrate(u0, mesh=mesh, dt=0.005, step=3)
    assert type(fused._state_dict["integrator"]).__name__ == "LoweredIntegrator"
    assert _rel(got, want) < 1e-12

    class LeansOnInternals(rbase.NonlinearFunc):
        def __call__(self, u_fft, f_mesh, u=None):
            return u_fft * f_mesh.no_such_table()
    bad = 0.05 * rop.Laplacian() + rop.NonlinearOperator(LeansOnInternals())
    with pytest.raises(NotImplementedError):
        reference_adapter.install(bad, strict=True).integrate(u0, mesh=mesh, dt=0.005, step=1)


def make_closure(base):
    class Closure(base.NonlinearFunc, torch.nn.Module):          # a learned closure: parameters inside a user core
        def __init__(self):
            torch.nn.Module.__init__(self)
            base.NonlinearFunc.__init__(self, True)
            self.a = torch.nn.Parameter(torch.tensor(0.7, dtype=torch.float64))
            self.b = torch.nn.Parameter(torch.tensor(-0.2, dtype=torch.float64))

        def __call__(self, u_fft, f_mesh, u=None):
            g = f_mesh.ifft(f_mesh.nabla_vector(1) * u_fft).real
            return f_mesh.fft(self.a * (-u ** 3) + self.b * (g * g).sum(dim=1, keepdim=True))
    return Closure


def _closure_run(pkg, base, mesh, etd, u0, w):
    core = make_closure(base)()
    op = 0.05 * pkg.Laplacian() + pkg.NonlinearOperator(core)
    op.set_integrator(etd.ETDRK2)
    x = u0.clone().requires_grad_(True)
    y = op.integrate(x, mesh=mesh, dt=0.005, step=3)
    (y * w).sum().backward()
    out = [y.detach(), x.grad, core.a.grad.clone(), core.b.grad.clone()]
    core.a.grad = core.b.grad = None
    op = 0.05 * pkg.Laplacian() + pkg.NonlinearOperator(core)     # parameters alone require grad
    op.set_integrator(etd.ETDRK2)
    (op.integrate(u0.clone(), mesh=mesh, dt=0.005, step=3) * w).sum().backward()
    return out + [core.a.grad.clone(), core.b.grad.clone()]


@pytest.mark.parametrize("mesh_info,shape", [([(0, 1, 32)], (2, 1, 32)), ([(0, 1, 16), (0, 1, 32)], (2, 1, 16, 32)),
                                             ([(0, 1, 8), (0, 1, 16), (0, 1, 8)], (2, 1, 8, 16, 8))], ids=["1d", "2d", "3d"])
def test_learned_closure_is_differentiable(mesh_info, shape):
    """Gradient mode hands a user core differentiable ``f_mesh.fft`` / ``.ifft`` and full spectra: gradients reach the
    initial field and the core's own parameters. Against the reference's autograd on the same source where importable,
    against a finite difference of the fused (no-grad) path always."""
    import torchfsm_b200 as fsm
    u0 = _u0(*shape)
    w = torch.randn(shape, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    mesh = fsm.MeshGrid(mesh_info, dtype=torch.float64)
    got = _closure_run(fsm, fsm, mesh, fsm.ETDRKIntegrator, u0, w)
    assert _rel(got[4], got[2]) < 1e-12 and _rel(got[5], got[3]) < 1e-12
    core = make_closure(fsm)()
    op = 0.05 * fsm.Laplacian() + fsm.NonlinearOperator(core)
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)

    def loss():
        with torch.no_grad():
            return float((op.integrate(u0.clone(), mesh=mesh, dt=0.005, step=3) * w).sum())
    eps = 1e-6
    with torch.no_grad():
        core.a += eps
        up = loss()
        core.a -= 2 * eps
        down = loss()
        core.a += eps
    assert type(op._state_dict["integrator"]).__name__ == "HostComposedStepper"
    assert abs((up - down) / (2 * eps) - float(got[2])) < 1e-6 * max(1.0, abs(float(got[2])))
    if os.path.isdir("/root/reference/torchfsm"):
        if "/root/reference" not in sys.path:
            sys.path.insert(0, "/root/reference")
        import torchfsm.operator as rop
        import torchfsm.operator._base as rbase
        import torchfsm.integrator as rint
        from torchfsm.mesh import MeshGrid as RefMesh
        want = _closure_run(rop, rbase, RefMesh(mesh_info, dtype=torch.float64), rint.ETDRKIntegrator, u0, w)
        for g, r in zip(got, want):
            assert _rel(g, r) < 1e-11


def test_custom_core_with_a_complex_symbol_on_a_2d_grid():
    """User core + odd-order linear term on a 2-D grid: the paired half spectra of unrolled.py; a core that reads the
    un-dealiased state is refused there (it would see the non-Hermitian Nyquist planes)."""
    import torchfsm_b200 as fsm
    Cubic, GradientSquared, _, _ = make_cores(fsm)
    mesh = fsm.MeshGrid([(0, 1, 16), (0, 1, 32)], dtype=torch.float64)
    u0 = _u0(2, 1, 16, 32)
    adv = 0.4 * fsm.SpatialDerivative(1, 1)
    custom = 0.05 * fsm.Laplacian() + adv + fsm.NonlinearOperator(Cubic())
    builtin = 0.05 * fsm.Laplacian() + adv + fsm.ImplicitSource(lambda u: -u ** 3)
    for op in (custom, builtin):
        op.set_integrator(fsm.SETDRKIntegrator.SETDRK2)
    got, want = (op.integrate(u0, mesh=mesh, dt=0.01, step=3) for op in (custom, builtin))
    assert type(custom._state_dict["integrator"]).__name__ == "PairedSpectrumStepper"
    assert _rel(got, want) < 1e-13
    with pytest.raises(NotImplementedError):
        (0.05 * fsm.Laplacian() + adv + fsm.NonlinearOperator(GradientSquared())).integrate(u0, mesh=mesh, dt=0.01, step=1)
