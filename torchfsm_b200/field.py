"""Initial-condition / forcing helpers used to build the synthetic inputs of the configs
(mirrors of ``torchfsm/field.py:8-61`` and ``:128-148``)."""
from typing import Optional

import torch

from .mesh import FourierMesh, MeshGrid
from .operator import Operator, Laplacian, ImplicitSource, ExplicitSource


def diffused_noise(mesh, diffusion_coef: float = 1.0, zero_centered: bool = True, unit_variance: bool = False,
                   unit_magnitude: bool = True, device=None, dtype=None, batch_size: int = 1, n_channel: int = 1,
                   generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """White noise diffused for one unit of time (one exact ETDRK0 step of diffusion_coef*Laplacian)."""
    if unit_magnitude and unit_variance:
        raise ValueError("unit_magnitude and unit_variance are mutually exclusive.")
    if device is None and isinstance(mesh, (FourierMesh, MeshGrid)):
        device = mesh.device
    if dtype is None and isinstance(mesh, (FourierMesh, MeshGrid)):
        dtype = mesh.dtype
    info = mesh.mesh_info if isinstance(mesh, (FourierMesh, MeshGrid)) else mesh
    shape = [batch_size, n_channel] + [m[2] for m in info]
    # drawn on the CPU so that the sample does not depend on the device
    u_0 = torch.randn(*shape, dtype=dtype, generator=generator).to(device)
    u_0 = (diffusion_coef * Laplacian()).integrate(u_0, dt=1, step=1, mesh=FourierMesh(info, device=device, dtype=dtype))
    dims = list(range(1, u_0.ndim))
    if zero_centered:
        u_0 = u_0 - u_0.mean(dim=dims, keepdim=True)
    if unit_variance:
        u_0 = u_0 / u_0.std(dim=dims, keepdim=True)
    if unit_magnitude:
        u_0 = u_0 / u_0.abs().amax(dim=dims, keepdim=True)
    return u_0


def kolm_force(x: torch.Tensor, drag_coef: float = -0.1, k: float = 4.0, length_scale: float = 1.0) -> Operator:
    """Kolmogorov forcing in vorticity form: drag*w - k cos(k l x)   (field.py:128-148)"""
    return drag_coef * ImplicitSource() - ExplicitSource(k * torch.cos(k * length_scale * x))
