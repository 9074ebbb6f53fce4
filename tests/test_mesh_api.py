"""FourierMesh / MeshGrid public surface (mesh.py:178-192, 244-266, 399-491): tables against the
reference's expressions and ``fft`` / ``ifft`` -- the reference's transform choke point -- running on the library's
passes (emulator build here) against ``torch.fft``."""
import os
import sys

import pytest
import torch

from product_util import build_emulator

MESHES = [[(0, 1, 16)], [(0, 1, 8), (0, 2, 16)], [(0, 1, 8), (0, 2, 16), (0, 3, 8)]]


@pytest.fixture(scope="module", autouse=True)
def emulator():
    from torchfsm_b200 import _cabi
    prev = _cabi._lib
    _cabi.use_library(build_emulator())
    yield
    _cabi._lib = prev


def _err(a, b):
    return float((a - b).abs().max())


@pytest.mark.parametrize("mesh_info", MESHES, ids=["1d", "2d", "3d"])
def test_transforms_match_torch_fft(mesh_info):
    import torchfsm_b200 as fsm
    torch.manual_seed(len(mesh_info))
    f = fsm.FourierMesh(mesh_info, dtype=torch.float64)
    shape = [m[2] for m in mesh_info]
    dims = list(range(-len(shape), 0))
    u = torch.randn(2, 3, *shape, dtype=torch.float64)
    spec = torch.randn(2, 3, *shape, dtype=torch.complex128)                 # not Hermitian
    assert _err(f.fft(u), torch.fft.fftn(u, dim=dims)) < 1e-12
    assert _err(f.fft(spec), torch.fft.fftn(spec, dim=dims)) < 1e-12
    assert _err(f.ifft(spec), torch.fft.ifftn(spec, dim=dims)) < 1e-13
    assert _err(f.ifft(f.fft(u)).real, u) < 1e-13 and float(f.ifft(f.fft(u)).imag.abs().max()) < 1e-13
    stacked = torch.randn(2, 2, 3, *shape, dtype=torch.float64)              # (B, d, C, N...) as in _convection.py:45-46
    assert _err(f.fft(stacked), torch.fft.fftn(stacked, dim=dims)) < 1e-12
    assert f.fft_dim == tuple(-(i + 1) for i in range(len(shape)))


@pytest.mark.parametrize("mesh_info", MESHES, ids=["1d", "2d", "3d"])
def test_tables_follow_the_reference_expressions(mesh_info):
    import torchfsm_b200 as fsm
    f = fsm.FourierMesh(mesh_info, dtype=torch.float64)
    d, shape = len(mesh_info), tuple(m[2] for m in mesh_info)
    assert len(f.bf) == len(f.f) == d
    for i, (a, b, n) in enumerate(mesh_info):
        assert torch.equal(f.f[i], torch.fft.fftfreq(n, (b - a) / n, dtype=torch.float64))
        assert torch.equal(f.bf[i], f.bf(i)) and f.bf[i].shape[i + 2] == n and f.bf[i].numel() == n
    assert torch.equal(f.f_x, f.f[0]) and torch.equal(f.bf_x, f.bf[0])
    with pytest.raises(ValueError):
        f.bf[d]
    assert f.bf_vector.shape == (1, d) + shape
    lap = sum((2j * torch.pi * f.bf[i]) ** 2 for i in range(d))
    assert torch.equal(f.laplacian(), lap)
    assert torch.equal(f.invert_laplacian(), torch.where(lap == 0, 1.0, 1 / lap))
    assert torch.equal(f.nabla_vector(3), (2j * torch.pi * f.bf_vector) ** 3)
    mask = f.low_pass_filter()
    assert mask.shape == (1, 1) + shape and set(mask.unique().tolist()) == {0.0, 1.0}
    kmax = f.low_pass_kmax(2 / 3)                                            # the box the kernels use == the mask
    box = torch.ones((1, 1) + shape, dtype=torch.float64)
    for i, n in enumerate(shape):
        idx = torch.arange(n)
        keep = (torch.where(idx <= n // 2, idx, n - idx) <= kmax[i]).to(torch.float64)
        view = [1] * (d + 2)
        view[i + 2] = n
        box = box * keep.reshape(view)
    assert torch.equal(mask, box)
    f.set_default_rel_freq_threshold(0.5)
    assert float(f.low_pass_filter().sum()) < float(mask.sum())
    assert float(f.abs_low_pass_filter(2).sum()) <= float(f.abs_low_pass_filter(3).sum())
    grid = fsm.MeshGrid(mesh_info, dtype=torch.float64)
    assert len(grid.meshs) == d and torch.equal(grid.meshs[0], grid.x)


def test_against_the_reference_mesh_when_importable():
    if not os.path.isdir("/root/reference/torchfsm"):
        pytest.skip("reference not present")
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    from torchfsm.mesh import FourierMesh as RefMesh
    import torchfsm_b200 as fsm
    for mesh_info in MESHES:
        f, r = fsm.FourierMesh(mesh_info, dtype=torch.float32), RefMesh(mesh_info, dtype=torch.float32)
        for name, args in (("low_pass_filter", ()), ("low_pass_filter", (0.5,)), ("abs_low_pass_filter", (3,)),
                           ("invert_laplacian", ()), ("invert_nabla", (2,)), ("nabla_vector", (1,)), ("grad", (0, 3)),
                           ("laplacian", ()), ("nabla", (4,))):
            assert torch.equal(getattr(f, name)(*args), getattr(r, name)(*args)), name
        assert torch.equal(f.bf_vector, r.bf_vector)
