"""Complex linear symbols on 2-D/3-D grids (torchfsm_b200/unrolled.py): properties of the paired half-spectrum state on
the host emulator build. Parity with the reference is in the ops fixtures (`*_complex_*`, `beta_plane*`, `*dispersion*`,
`*advection*`; tests/test_emu_ops.py, tests/test_gpu_parity.py) and in tests/test_reference_adapter.py."""
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


@pytest.mark.parametrize("mesh_info", [[(0, 6.28, 8), (0, 6.28, 16)], [(0, 1, 8), (0, 1, 16), (0, 1, 8)]], ids=["2d", "3d"])
def test_pair_holds_any_full_spectrum(mesh_info):
    import torchfsm_b200 as fsm
    from torchfsm_b200.unrolled import PairedSpectrumStepper
    torch.manual_seed(2)
    mesh = fsm.MeshGrid(mesh_info, dtype=torch.float64)
    op = 0.01 * fsm.Laplacian() + 0.5 * fsm.SpatialDerivative(0, 1) - 0.1 * fsm.SpatialDerivative(len(mesh_info) - 1, 3)
    shape = [m[2] for m in mesh_info]
    dims = list(range(2, 2 + len(shape)))
    u = torch.randn(2, 1, *shape, dtype=torch.float64)
    out = op.integrate(u, mesh=mesh, dt=0.01, step=2)
    st = op._state_dict["integrator"]
    assert isinstance(st, PairedSpectrumStepper)
    full = torch.randn(2, 1, *shape, dtype=torch.complex128)              # not Hermitian
    pair = st.full_to_half(full)
    assert float((st.half_to_full(pair) - full).abs().max()) == 0.0       # lossless both ways
    # the physical field of a pair is the real part of the inverse transform of the full spectrum (_base.py:747-751)
    assert float((st.c2r(pair) - torch.fft.ifftn(full, dim=dims).real).abs().max()) < 1e-13
    # a real field maps to two equal members, and stepping the full-spectrum protocol equals stepping the pair
    x = st.r2c(u)
    assert float((x[0] - x[1]).abs().max()) == 0.0
    f = torch.fft.fftn(u, dim=dims)
    for _ in range(2):
        f = st.step(f)
    assert float((torch.fft.ifftn(f, dim=dims).real - out).abs().max()) < 1e-13
    # after a step the members differ exactly on the Nyquist planes of the odd-order axes
    st.step_half(x, 1)
    diff = (st.half_to_full(torch.stack([x[0], x[0]])) - st.half_to_full(x)).abs()
    assert float(diff.max()) > 1e-6
    interior = diff
    for a, n in enumerate(shape):
        interior = interior.index_select(2 + a, torch.tensor([i for i in range(n) if i != n // 2]))
    assert float(interior.max()) < 1e-13


def test_linear_advection_translates_the_field():
    """u_t = c u_x moves a band-limited field by c t exactly (ETDRK0 is exact for linear operators)."""
    import math
    import torchfsm_b200 as fsm
    mesh = fsm.MeshGrid([(0, 1, 32), (0, 1, 16)], dtype=torch.float64)
    x, y = mesh.bc_mesh_grid()
    f = lambda xx, yy: torch.sin(2 * math.pi * xx) * torch.cos(4 * math.pi * yy) + 0.3 * torch.cos(6 * math.pi * xx)
    c, t = 0.7, 0.25
    out = (c * fsm.SpatialDerivative(0, 1)).integrate((f(x, y) + 0 * x).contiguous(), mesh=mesh, dt=t / 5, step=5)
    assert float((out - f(x + c * t, y)).abs().max()) < 1e-12


def test_restart_from_a_spectral_checkpoint_is_exact():
    """Manual checkpoint / resume: ``return_in_fourier`` then ``u_0_fft``. With a complex symbol the
    checkpoint is the reference's non-Hermitian full spectrum and the pair is rebuilt from it without loss."""
    import torchfsm_b200 as fsm
    torch.manual_seed(4)
    mesh = fsm.MeshGrid([(0, 6.28, 16), (0, 6.28, 16)], dtype=torch.float64)
    u0 = 0.5 * torch.randn(2, 1, 16, 16, dtype=torch.float64)
    for op in (0.01 * fsm.Laplacian() - fsm.VorticityConvection() + 0.5 * fsm.SpatialDerivative(0, 1),
               0.01 * fsm.Laplacian() - fsm.VorticityConvection()):
        op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
        straight = op.integrate(u0, mesh=mesh, dt=0.01, step=6)
        ckpt = op.integrate(u0, mesh=mesh, dt=0.01, step=3, return_in_fourier=True)
        resumed = op.integrate(u_0_fft=ckpt, mesh=mesh, dt=0.01, step=3)
        assert float((resumed - straight).abs().max()) < 1e-14
