"""PDE presets: the operator sums of the reference's ``torchfsm/pde.py`` that lie on the hot path."""
from typing import Optional

from .operator import (Operator, Convection, Laplacian, Biharmonic, KSConvection, VorticityConvection,
                       NSPressureConvection)


def Burgers(nu: float) -> Operator:
    """du/dt = -u.grad(u) + nu lap(u)   (pde.py:13-26)"""
    return nu * Laplacian() - Convection()


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
