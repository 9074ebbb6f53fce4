"""Drop-in test: a GENUINE torchfsm operator, its own unmodified ``integrate`` loop
(operator/_base.py:676-751), stepping through the fused kernels via ``torchfsm_b200.reference_adapter``.
CPU: the host-emulator build of the kernels; skipped where the reference cannot be imported (it never is on the
GPU box unless baseline/_ref travelled there). The GPU variant lives in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "torchfsm")):
            if cand not in sys.path:
                sys.path.append(cand)
            try:
                import torchfsm
                return torchfsm
            except Exception:
                continue
    return None


def reference_cases(torchfsm, device, dtype):
    """(name, operator, mesh, u0, dt, steps) built ONLY with the reference's public API."""
    from torchfsm.mesh import MeshGrid
    from torchfsm.pde import Burgers, NavierStokesVorticity, NavierStokes, KuramotoSivashinskyHighDim
    from torchfsm.field import kolm_force
    from torchfsm.integrator import ETDRKIntegrator, SETDRKIntegrator
    g = torch.Generator().manual_seed(0)
    out = []
    m1 = MeshGrid([(0, 1, 128)], device=device, dtype=dtype)                       # C1 (README.md:30-53)
    x = m1.bc_mesh_grid()
    out.append(("c1_burgers1d", Burgers(0.01), m1, torch.sin(2 * torch.pi * x) + 0.5, 0.01, 20))
    m3 = MeshGrid([(0, 2 * np.pi, 32)] * 2, device=device, dtype=dtype)            # C3 shape
    _, y = m3.bc_mesh_grid()
    op3 = NavierStokesVorticity(Re=100, force=kolm_force(y))
    op3.set_integrator(ETDRKIntegrator.ETDRK2)
    out.append(("c3_ns2d", op3, m3, torch.randn(3, 1, 32, 32, generator=g, dtype=dtype).to(device), 0.01, 5))
    m2 = MeshGrid([(0, 60, 32)] * 2, device=device, dtype=dtype)                   # C2 shape
    out.append(("c2_ks2d", KuramotoSivashinskyHighDim(), m2,
                torch.randn(4, 1, 32, 32, generator=g, dtype=dtype).to(device), 0.1, 3))
    m5 = MeshGrid([(0, 2 * np.pi, 16)] * 3, device=device, dtype=dtype)            # C5 shape
    op5 = NavierStokes(Re=100)
    op5.set_integrator(SETDRKIntegrator.SETDRK4)
    u5 = 0.3 * torch.randn(1, 3, 16, 16, 16, generator=g, dtype=dtype).to(device)
    out.append(("c5_ns3d", op5, m5, u5, 0.01, 3))
    # around the path: what the adapter translates through this package's own lowering
    from torchfsm.operator import (Laplacian, Convection, ConservativeConvection, ImplicitSource, ExplicitSource,
                                   NSPressureConvection)
    from torchfsm.integrator import RKIntegrator
    m4 = MeshGrid([(0, 1, 32), (0, 1, 16)], device=device, dtype=dtype)
    u4 = 0.5 * torch.randn(2, 2, 32, 16, generator=g, dtype=dtype).to(device)
    out.append(("conservative_convection2d", 0.01 * Laplacian() - ConservativeConvection(), m4, u4, 1e-3, 3))
    out.append(("allen_cahn2d", 0.05 * Laplacian() + ImplicitSource(lambda u: u - u ** 3), m3,
                torch.randn(2, 1, 32, 32, generator=g, dtype=dtype).to(device), 0.01, 3))
    xx, yy, zz = m5.bc_mesh_grid()
    body = torch.cat([0.3 * torch.sin(yy) * torch.cos(2 * zz) + 0 * xx, 0.2 * torch.cos(xx + zz) + 0 * yy,
                      0.1 * torch.sin(2 * xx) * torch.sin(yy) + 0 * zz], dim=1)
    opf = NSPressureConvection(ExplicitSource(body)) + 0.01 * Laplacian()
    opf.set_integrator(ETDRKIntegrator.ETDRK2)
    out.append(("ns3d_body_force", opf, m5, u5, 0.005, 3))
    opr = 0.01 * Laplacian() - Convection()
    opr.set_integrator(RKIntegrator.Dorpi45)
    out.append(("burgers2d_dorpi45", opr, m4, u4, 2e-4, 3))
    # complex linear symbol on a 2-D grid: the reference's loop hands its (non-Hermitian) full spectrum to every step
    from torchfsm.operator import SpatialDerivative, VorticityConvection
    opb = 0.01 * Laplacian() - VorticityConvection() + 0.5 * SpatialDerivative(0, 1)
    opb.set_integrator(ETDRKIntegrator.ETDRK2)
    out.append(("beta_plane2d", opb, m3, torch.randn(2, 1, 32, 32, generator=g, dtype=dtype).to(device), 0.01, 4))
    nu = torch.tensor([0.01, 0.03], dtype=dtype, device=device).reshape(2, 1, 1, 1)
    out.append(("burgers2d_batched_nu", nu * Laplacian() - Convection(), m4, u4, 1e-3, 3))
    return out


def run_dropin(torchfsm, device, dtype, tol, only=None):
    from torchfsm_b200 import reference_adapter
    import copy
    for name, op, mesh, u0, dt, steps in reference_cases(torchfsm, device, dtype):
        if only is not None and name not in only:
            continue
        want = copy.deepcopy(op).integrate(u0, mesh=mesh, dt=dt, step=steps)        # stock torch path
        fused = reference_adapter.install(copy.deepcopy(op), strict=True)
        got = fused.integrate(u0, mesh=mesh, dt=dt, step=steps)                    # the reference's OWN loop
        assert type(fused._state_dict["integrator"]).__name__ == "LoweredIntegrator", name
        err = float((got - want).norm() / want.norm())
        assert err <= tol * steps, (name, err)
        # the reference's recorder protocol on full-spectrum frames
        from torchfsm.traj_recorder import AutoRecorder
        traj = fused.integrate(u0, dt=dt, step=2, trajectory_recorder=AutoRecorder())
        assert traj.shape[1] == 3 and float((traj[:, 0] - u0).abs().max()) < 1e-5


def test_reference_operator_runs_on_the_fused_kernels_emulator():
    torchfsm = _reference()
    if torchfsm is None:
        pytest.skip("reference not importable here")
    from product_util import build_emulator
    from torchfsm_b200 import _cabi
    _cabi.use_library(build_emulator())
    run_dropin(torchfsm, "cpu", torch.float64, 1e-12)
    # fp32 on the CPU: the config-shaped cases (the GPU run covers every case in both precisions)
    run_dropin(torchfsm, "cpu", torch.float32, 1e-5, only=("c1_burgers1d", "c3_ns2d", "c2_ks2d", "c5_ns3d", "beta_plane2d"))


def test_unsupported_reference_operators_fall_back_to_torch():
    torchfsm = _reference()
    if torchfsm is None:
        pytest.skip("reference not importable here")
    from product_util import build_emulator
    from torchfsm_b200 import _cabi, reference_adapter
    from torchfsm.mesh import MeshGrid
    from torchfsm.pde import Burgers
    _cabi.use_library(build_emulator())
    mesh = MeshGrid([(0, 1, 12), (0, 1, 12)], dtype=torch.float64)                  # 12 points: not a power of two
    u0 = torch.randn(1, 2, 12, 12, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    op = reference_adapter.install(Burgers(0.01))
    out = op.integrate(u0, mesh=mesh, dt=1e-3, step=2)
    assert type(op._state_dict["integrator"]).__name__ != "LoweredIntegrator" and torch.isfinite(out).all()
    with pytest.raises(NotImplementedError):
        reference_adapter.install(Burgers(0.01), strict=True).integrate(u0, mesh=mesh, dt=1e-3, step=1)


def test_truncated_fourier_series_equals_the_reference_for_a_seed():
    """The IC helper of the batched-parameter tutorial (field.py:62-125): same draws from torch's generator, inverse
    transform on the library (emulator build here) -> the same field as the reference."""
    torchfsm = _reference()
    if torchfsm is None:
        pytest.skip("reference not importable")
    import torchfsm.field      # noqa: F401  (the reference's package does not import its submodules)
    import torchfsm.mesh       # noqa: F401
    from product_util import build_emulator
    from torchfsm_b200 import _cabi
    import torchfsm_b200 as fsm
    prev = (_cabi._lib, _cabi._lib_path)          # another module's fixture may own the loaded library (xdist)
    _cabi.use_library(build_emulator())
    try:
        for info, kw in [([(0, 1.0, 128)], dict(batch_size=3, freq_threshold=2)),
                         ([(0, 2.0, 32), (0, 1.0, 16)], dict(batch_size=2, n_channel=2, freq_threshold=5, unit_magnitude=False,
                                                            unit_variance=True))]:
            torch.manual_seed(5)
            want = torchfsm.field.truncated_fourier_series(torchfsm.mesh.MeshGrid(info, dtype=torch.float64), **kw)
            torch.manual_seed(5)
            got = fsm.field.truncated_fourier_series(fsm.MeshGrid(info, dtype=torch.float64), **kw)
            assert got.shape == want.shape and float((got - want).abs().max()) < 1e-12
    finally:
        _cabi._lib, _cabi._lib_path = prev
