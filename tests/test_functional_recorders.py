"""Function forms of the operators (torchfsm_b200/functional.py), the disk / per-sample recorders and ``wave_1d``, on the
host emulator build; against the unmodified reference where it is importable (authoring container), against this
package's operator classes otherwise."""
import os
import sys

import numpy as np
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


def _reference():
    if os.path.isdir("/root/reference/torchfsm") and "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    try:
        import torchfsm.functional  # noqa: F401
        import torchfsm
        return torchfsm
    except Exception:
        return None


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def _smooth(*shape):
    import torchfsm_b200 as fsm
    g = torch.Generator().manual_seed(sum(shape))
    u = torch.randn(*shape, dtype=torch.float64, generator=g)
    mesh = fsm.MeshGrid([(0, 1, s) for s in shape[2:]], dtype=torch.float64)
    return (0.002 * fsm.Laplacian()).integrate(u, mesh=mesh, dt=1.0, step=1)


MI2, MI3 = [(0, 6.28, 16), (0, 3.0, 32)], [(0, 1, 8), (0, 2, 16), (0, 6.28, 8)]
FORMS = [("biharmonic", (2, 1, 16, 32), MI2, {}, "Biharmonic", ()), ("laplacian", (2, 2, 16, 32), MI2, {}, "Laplacian", ()),
         ("grad", (2, 1, 16, 32), MI2, {}, "Grad", ()), ("div", (2, 2, 16, 32), MI2, {}, "Div", ()),
         ("curl", (2, 3, 8, 16, 8), MI3, {}, "Curl", ()), ("convection", (2, 2, 16, 32), MI2, {}, "Convection", ()),
         ("conservative_convection", (2, 3, 8, 16, 8), MI3, {}, "ConservativeConvection", ()),
         ("ks_convection", (2, 1, 16, 32), MI2, {"remove_mean": False}, "KSConvection", (False,)),
         ("vorticity_convection", (2, 1, 16, 32), MI2, {}, "VorticityConvection", ()),
         ("vorticity2velocity", (2, 1, 16, 32), MI2, {}, "Vorticity2Velocity", ()),
         ("velocity2pressure", (2, 3, 8, 16, 8), MI3, {}, "Velocity2Pressure", ()),
         ("vorticity2pressure", (2, 1, 16, 32), MI2, {}, "Vorticity2Pressure", ())]


@pytest.mark.parametrize("form", FORMS, ids=[f[0] for f in FORMS])
def test_function_forms(form):
    import torchfsm_b200 as fsm
    import torchfsm_b200.functional as F
    name, shape, mi, kw, cls, args = form
    u = _smooth(*shape)
    mesh = fsm.MeshGrid(mi, dtype=torch.float64)
    got = getattr(F, name)(u.clone(), mesh=mesh, **kw)
    assert _rel(got, getattr(fsm, cls)(*args)(u.clone(), mesh=mesh)) == 0.0
    key = {"vorticity2velocity": "vorticity_fft", "vorticity2pressure": "vorticity_fft",
           "velocity2pressure": "velocity_fft"}.get(name, "u_fft")
    spectral = getattr(F, name)(**{key: torch.fft.fftn(u, dim=list(range(2, u.dim())))}, mesh=mesh, **kw)
    assert _rel(spectral, got) < 1e-13
    ref = _reference()
    if ref is not None:
        from torchfsm.mesh import MeshGrid
        want = getattr(ref.functional, name)(u.clone(), mesh=MeshGrid(mi, dtype=torch.float64), **kw)
        assert _rel(got, want) < 1e-12


def test_spatial_derivative_and_forced_pressure_forms():
    import torchfsm_b200 as fsm
    import torchfsm_b200.functional as F
    u = _smooth(2, 1, 16, 32)
    mesh = fsm.MeshGrid(MI2, dtype=torch.float64)
    got = F.spatial_derivative(1, 3, u.clone(), mesh=mesh)
    assert _rel(got, fsm.SpatialDerivative(1, 3)(u.clone(), mesh=mesh)) == 0.0
    p = F.vorticity2pressure(u.clone(), mesh=mesh, external_force=-0.1 * fsm.ImplicitSource())
    ref = _reference()
    if ref is not None:
        from torchfsm.mesh import MeshGrid
        import torchfsm.operator as rop
        rm = MeshGrid(MI2, dtype=torch.float64)
        assert _rel(got, ref.functional.spatial_derivative(1, 3, u.clone(), mesh=rm)) < 1e-12
        assert _rel(p, ref.functional.vorticity2pressure(u.clone(), mesh=rm, external_force=-0.1 * rop.ImplicitSource())) < 1e-12


def test_disk_recorder_writes_the_trajectory_in_chunks(tmp_path):
    import torchfsm_b200 as fsm
    mesh = fsm.MeshGrid([(0, 1, 32)], dtype=torch.float64)
    u0 = _smooth(2, 1, 32)
    op = fsm.pde.Burgers(0.01)
    whole = op.integrate(u0, mesh=mesh, dt=0.01, step=6, trajectory_recorder=fsm.AutoRecorder())        # (B, 7, C, N)
    for fmt in ("torch", "numpy"):
        d = str(tmp_path / fmt) + os.sep
        rec = fsm.DiskRecorder(cache_dir=d, cache_freq=3, save_format=fmt)
        assert op.integrate(u0, mesh=mesh, dt=0.01, step=6, trajectory_recorder=rec) is None
        assert len(rec.files) == 3                      # 3 + 3 + the last frame flushed at the end
        parts = [torch.from_numpy(np.load(f)) if fmt == "numpy" else torch.load(f) for f in rec.files]
        assert [p.shape[1] for p in parts] == [3, 3, 1]
        assert float((torch.cat(parts, dim=1) - whole).abs().max()) == 0.0
    with pytest.raises(ValueError):
        fsm.DiskRecorder(save_format="hdf5")


def test_random_batch_wise_recorder():
    import torchfsm_b200 as fsm
    mesh = fsm.MeshGrid([(0, 1, 32)], dtype=torch.float64)
    u0 = _smooth(4, 1, 32)
    op = fsm.pde.Burgers(0.01)
    whole = op.integrate(u0, mesh=mesh, dt=0.01, step=20, trajectory_recorder=fsm.AutoRecorder())
    np.random.seed(7)
    rec = fsm.RandomBatchWisedRecorder(simulation_steps=20, recorder_interval=3, n_recorded_frames=3)
    traj = op.integrate(u0, mesh=mesh, dt=0.01, step=20, trajectory_recorder=rec)
    assert traj.shape == (4, 3, 1, 32)
    ids = rec._recorded_frame_id
    assert ids.shape == (4, 3) and (np.diff(ids, axis=1) == 3).all() and len(set(ids[:, 0])) > 1
    for b in range(4):
        for j in range(3):
            assert float((traj[b, j] - whole[b, ids[b, j]]).abs().max()) < 1e-13
    ref = _reference()
    if ref is not None:                                  # same numpy seed -> same frames as the reference's recorder
        from torchfsm.mesh import MeshGrid
        from torchfsm.pde import Burgers
        from torchfsm.traj_recorder import RandomBatchWisedRecorder
        np.random.seed(7)
        want = Burgers(0.01).integrate(u0, mesh=MeshGrid([(0, 1, 32)], dtype=torch.float64), dt=0.01, step=20,
                                       trajectory_recorder=RandomBatchWisedRecorder(20, 3, 3))
        assert _rel(traj, want) < 1e-12


def test_wave_1d_draws_like_the_reference():
    import torchfsm_b200 as fsm
    mesh = fsm.MeshGrid([(0, 1, 64)], dtype=torch.float64)
    x = mesh.bc_mesh_grid()
    torch.manual_seed(11)
    y = fsm.field.wave_1d(x, zero_mean=False)
    assert y.shape == x.shape and torch.isfinite(y).all() and float(y.std()) > 0.1
    torch.manual_seed(11)
    yb = fsm.field.wave_1d(x.expand(3, 1, 64).contiguous(), batched=True, zero_mean=True)
    assert yb.shape == (3, 1, 64) and float((yb[0] - yb[1]).abs().max()) > 1e-3
    ref = _reference()
    if ref is not None:
        from torchfsm.field import wave_1d
        torch.manual_seed(11)
        assert float((wave_1d(x, zero_mean=False) - y).abs().max()) < 1e-14
