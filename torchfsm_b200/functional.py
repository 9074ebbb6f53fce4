"""Function forms of the operators (mirror of ``torchfsm/functional.py:9-360``): ``grad(u, mesh=mesh)`` is
``Grad()(u, mesh=mesh)`` and so on -- one evaluation on the library's passes (``fsm_r2c`` -> ``fsm_spectral_map`` or the
fused right-hand side -> ``fsm_c2r``). Same argument names as the reference; every function takes the field in physical
space (``u``) or as a full spectrum (``u_fft``)."""
from typing import Optional

import torch

from .operator import (Biharmonic, ConservativeConvection, Convection, Curl, Div, Grad, KSConvection, Laplacian,
                       OperatorLike, SpatialDerivative, Velocity2Pressure, Vorticity2Pressure, Vorticity2Velocity,
                       VorticityConvection)


def biharmonic(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return Biharmonic()(u=u, u_fft=u_fft, mesh=mesh)


def conservative_convection(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return ConservativeConvection()(u=u, u_fft=u_fft, mesh=mesh)


def convection(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return Convection()(u=u, u_fft=u_fft, mesh=mesh)


def curl(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return Curl()(u=u, u_fft=u_fft, mesh=mesh)


def div(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return Div()(u=u, u_fft=u_fft, mesh=mesh)


def grad(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return Grad()(u=u, u_fft=u_fft, mesh=mesh)


def laplacian(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return Laplacian()(u=u, u_fft=u_fft, mesh=mesh)


def spatial_derivative(dim_index: int, order: int, u: Optional[torch.Tensor] = None,
                       u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return SpatialDerivative(dim_index, order)(u=u, u_fft=u_fft, mesh=mesh)


def ks_convection(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None,
                  remove_mean: bool = True) -> torch.Tensor:
    return KSConvection(remove_mean)(u=u, u_fft=u_fft, mesh=mesh)


def vorticity_convection(u: Optional[torch.Tensor] = None, u_fft: Optional[torch.Tensor] = None, mesh=None) -> torch.Tensor:
    return VorticityConvection()(u=u, u_fft=u_fft, mesh=mesh)


def vorticity2velocity(vorticity: Optional[torch.Tensor] = None, vorticity_fft: Optional[torch.Tensor] = None,
                       mesh=None) -> torch.Tensor:
    return Vorticity2Velocity()(u=vorticity, u_fft=vorticity_fft, mesh=mesh)


def velocity2pressure(velocity: Optional[torch.Tensor] = None, velocity_fft: Optional[torch.Tensor] = None, mesh=None,
                      external_force: Optional[OperatorLike] = None) -> torch.Tensor:
    """functional.py:293-324: convection -> minus force -> divergence -> Poisson solve; here one evaluation of the
    convection program and one point-wise map (``OperatorLike._eval_composite``)."""
    return Velocity2Pressure(external_force)(u=velocity, u_fft=velocity_fft, mesh=mesh)


def vorticity2pressure(vorticity: Optional[torch.Tensor] = None, vorticity_fft: Optional[torch.Tensor] = None, mesh=None,
                       external_force: Optional[OperatorLike] = None) -> torch.Tensor:
    """functional.py:327-360."""
    return Vorticity2Pressure(external_force)(u=vorticity, u_fft=vorticity_fft, mesh=mesh)
