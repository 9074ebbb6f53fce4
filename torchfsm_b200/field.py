"""Initial-condition / forcing helpers used to build the synthetic inputs of the configs
(mirrors of ``torchfsm/field.py:8-61`` and ``:128-148``)."""
from typing import Optional

import torch

from .mesh import FourierMesh, MeshGrid
from . import _cabi
from .operator import Operator, Laplacian, ImplicitSource, ExplicitSource, FusedStepper


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


def truncated_fourier_series(mesh, freq_threshold: int = 5, amplitude_range=(-1.0, 1.0), angle_range=(0.0, 2.0 * torch.pi),
                             zero_centered: bool = True, unit_variance: bool = False, unit_magnitude: bool = True,
                             device=None, dtype=None, batch_size: int = 1, n_channel: int = 1) -> torch.Tensor:
    """Random field with a truncated Fourier spectrum (field.py:62-125): random magnitudes and phases on the modes with
    |f_i| <= freq_threshold, the real part of the inverse transform (Hermitian projection + C2R on the library), then the
    same normalisations. Draws from torch's global generator in the reference's order, so a seed gives the same field."""
    if unit_magnitude and unit_variance:
        raise ValueError("Cannot set both unit_magnitude and unit_variance to True.")
    if device is None and isinstance(mesh, (FourierMesh, MeshGrid)):
        device = mesh.device
    if dtype is None and isinstance(mesh, (FourierMesh, MeshGrid)):
        dtype = mesh.dtype
    f_mesh = mesh if isinstance(mesh, FourierMesh) else FourierMesh(mesh, device=device, dtype=dtype)
    shape = [batch_size, n_channel] + list(f_mesh.shape)
    magnitude = torch.rand(*shape, device=device, dtype=dtype) * (amplitude_range[1] - amplitude_range[0]) + amplitude_range[0]
    angle = torch.rand(*shape, device=device, dtype=dtype) * (angle_range[1] - angle_range[0]) + angle_range[0]
    mask = torch.ones(shape[2:], device=f_mesh.device, dtype=f_mesh.dtype)
    for i in range(f_mesh.n_dim):                                       # mesh.py:463-479 (absolute frequency threshold)
        mask = mask * torch.where(f_mesh.bf(i)[0, 0].abs() > freq_threshold, 0, 1)
    noise_hat = (magnitude * torch.exp(1j * angle) * mask).to(f_mesh.device)
    st = FusedStepper(f_mesh, batch_size, n_channel, _cabi.PROG_LINEAR, "RK4", 1.0, None, 0.0, None,
                      [n // 2 for n in f_mesh.shape], True, {})
    noise = st.c2r(st.full_to_half(noise_hat))
    dims = list(range(1, noise.ndim))
    if zero_centered:
        noise = noise - noise.mean(dim=dims, keepdim=True)
    if unit_variance:
        noise = noise / noise.std(dim=dims, keepdim=True)
    if unit_magnitude:
        noise = noise / noise.abs().amax(dim=dims, keepdim=True)
    return noise


def wave_1d(x: torch.Tensor, min_k: int = 1, max_k: int = 5, min_amplitude: float = 0.5, max_amplitude: float = 1.0,
            n_polynomial: int = 5, zero_mean: bool = False, mean_shift_coef=0.3, batched: bool = False) -> torch.Tensor:
    """Sum of ``n_polynomial`` sines with random integer wavenumbers, amplitudes and phases on the coordinate field ``x``
    (field.py:151-207); draws from torch's global generator in the reference's order, so a seed gives the same field."""
    x_new = x / x.max() * torch.pi * 2
    y = torch.zeros_like(x)
    if not batched:
        x_new, y = x_new.unsqueeze(0), y.unsqueeze(0)
    batch = x_new.shape[0]
    shape = [batch, n_polynomial] + [1] * (x_new.dim() - 2)
    k = torch.randint(min_k, max_k + 1, shape, device=x.device, dtype=x.dtype)
    amplitude = torch.rand(*shape, device=x.device, dtype=x.dtype) * (max_amplitude - min_amplitude) + min_amplitude
    shift = torch.rand(*shape, device=x.device, dtype=x.dtype) * torch.pi * 2
    for i in range(n_polynomial):
        y = y + amplitude[:, i:i + 1] * torch.sin(k[:, i:i + 1] * (x_new + shift[:, i:i + 1]))
    if not zero_mean:
        value_shift = torch.rand([batch] + [1] * (x_new.dim() - 1), device=x.device, dtype=x.dtype)
        y = y + ((value_shift - 0.5) * 2 * (max_amplitude - min_amplitude) * mean_shift_coef + min_amplitude)
    return y if batched else y.squeeze(0)
