"""torchfsm_b200 — B200-native drop-in for ONE path of qiauil/torchfsm: the per-step
pseudo-spectral update behind ``Operator.integrate(u_0, mesh, dt, step)``.

Host code is Python over PyTorch tensors (device memory, streams); the work is done by
hand-written sm_100a CUDA kernels reached through a C ABI (``include/fsm_b200.h``,
``torchfsm_b200/libfsm_b200.so``). There is no torch/CPU fallback on this path.
"""
from .mesh import MeshGrid, FourierMesh  # noqa: F401
from .integrator import ETDRKIntegrator, SETDRKIntegrator, RKIntegrator  # noqa: F401
from .operator import (Operator, LinearOperator, NonlinearOperator, Laplacian, Biharmonic,  # noqa: F401
                       SpatialDerivative, ImplicitSource, ExplicitSource, Convection, KSConvection,
                       VorticityConvection, NSPressureConvection, FusedStepper, Grad, Div, Curl,
                       Vorticity2Velocity, Vorticity2Pressure, Velocity2Pressure, ConservativeConvection,
                       run_operators, HostComposedStepper, DynamicForceStepper, LinearCoef, NonlinearFunc,
                       CoreGenerator)
from .traj_recorder import (AutoRecorder, CPURecorder, DiskRecorder, RandomBatchWisedRecorder,  # noqa: F401
                            IntervalController)
from . import pde, field, functional  # noqa: F401

__version__ = "0.1.0"
