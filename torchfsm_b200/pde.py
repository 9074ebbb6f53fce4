"""PDE presets: the operator sums of the reference's ``torchfsm/pde.py`` that lie on the hot path."""
from typing import Optional

from .operator import (Operator, Convection, Laplacian, Biharmonic, KSConvection, VorticityConvection,
                       NSPressureConvection, SpatialDerivative)


def Burgers(nu: float) -> Operator:
    """du/dt = -u.grad(u) + nu lap(u)   (pde.py:13-26)"""
    return nu * Laplacian() - Convection()


def KuramotoSivashinsky() -> Operator:
    """1-D KS: dphi/dt = -phi_xx - phi_xxxx - phi phi_x   (pde.py:27-39)"""
    ks_eqn = -Laplacian() - Biharmonic() - Convection()
    ks_eqn.register_additional_check(lambda dim_value, dim_mesh: dim_value == 1 and dim_mesh == 1)
    return ks_eqn


def KortewegDeVries(dispersion_coef=1, convection_coef: float = 6.0) -> Operator:
    """dphi/dt = -c1 phi_xxx + c2 phi phi_x   (pde.py:51-64); the dispersion makes exp(L dt) complex"""
    return -dispersion_coef * SpatialDerivative(0, 3) + convection_coef * Convection()


def KuramotoSivashinskyHighDim() -> Operator:
    """dphi/dt = -lap(phi) - lap^2(phi) - 1/2 |grad phi|^2   (pde.py:41-49)"""
    return -Laplacian() - Biharmonic() - KSConvection()


def NavierStokesVorticity(Re: float, force: Optional[Operator] = None) -> Operator:
    """dw/dt + (u.grad) w = 1/Re lap(w) + curl f   (pde.py:66-83)"""
    ns_vorticity = -VorticityConvection() + 1 / Re * Laplacian()
    if force is not None:
        ns_vorticity += force
    return ns_vorticity


def NavierStokes(Re: float, force: Optional[Operator] = None) -> Operator:
    """Velocity-pressure form with the pressure projected out   (pde.py:85-100)"""
    return NSPressureConvection(force) + 1 / Re * Laplacian()
