"""Gradient mode on the GPU (CUDA library): the same cases as tests/test_autograd_nonlinear.py, against the gradients the
reference produced under its own autograd (tests/golden_grad)."""
import pytest
import torch

from grad_checks import check_grad_case
from grad_util import grad_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("name", grad_names())
def test_gpu_gradients_match_reference_autograd(name, dtype):
    from torchfsm_b200 import _cabi
    assert not _cabi.is_emulator()
    check_grad_case(name, "cuda", dtype)


def test_gpu_gradient_through_c3_shaped_step():
    """NS vorticity + Kolmogorov forcing, ETDRK2 (the C3 operator) at 256^2 x 2: the gradient of a quadratic loss after
    two steps against a central finite difference along one random direction."""
    import numpy as np
    import torchfsm_b200 as fsm
    dev = torch.device("cuda", 0)
    mesh = fsm.MeshGrid([(0, 2 * np.pi, 256)] * 2, device=dev, dtype=torch.float64)
    op = fsm.pde.NavierStokesVorticity(Re=100, force=fsm.field.kolm_force(mesh.bc_mesh_grid()[1]))
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    g = torch.Generator().manual_seed(5)
    u0 = fsm.field.diffused_noise(mesh, batch_size=2, generator=g)
    v = torch.randn(u0.shape, dtype=torch.float64, generator=g).to(dev)
    v = (0.01 * fsm.Laplacian()).integrate(v, mesh=mesh, dt=1.0, step=1)

    def loss(u):
        return 0.5 * (op.integrate(u, mesh=mesh, dt=0.01, step=2) ** 2).sum()
    x = u0.clone().requires_grad_(True)
    loss(x).backward()
    eps = 1e-5
    fd = float(loss(u0 + eps * v) - loss(u0 - eps * v)) / (2 * eps)
    an = float((x.grad * v).sum())
    assert abs(fd - an) < 1e-6 * max(1.0, abs(an)), (fd, an)
