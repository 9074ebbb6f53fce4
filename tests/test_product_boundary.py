"""Boundary hygiene: the product never touches the oracle or the emulator, the C-ABI library
exports what include/fsm_b200.h declares, and a missing CUDA library is a loud error."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "torchfsm_b200")


def _product_sources():
    for base, _, files in os.walk(PKG):
        if "_build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                yield os.path.join(base, f)


def test_product_does_not_import_oracle_or_emulator():
    for path in _product_sources():
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), path
        assert "tests.emu" not in src and "libfsm_emu" not in src, path
        assert "/root/reference" not in src, path


def test_header_symbols_are_exported_by_the_cuda_library():
    lib_path = os.path.join(PKG, "libfsm_b200.so")
    if not os.path.exists(lib_path):
        if shutil.which("nvcc") is None:
            pytest.skip("nvcc not available and libfsm_b200.so not built")
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(PKG, "csrc"), "cuda"], check=True)
    header = open(os.path.join(ROOT, "include", "fsm_b200.h")).read()
    declared = set(re.findall(r"\b(fsm_[a-z0-9_]+)\s*\(", header))
    assert {"fsm_plan_create", "fsm_step", "fsm_r2c", "fsm_c2r", "fsm_rhs"} <= declared
    lib = ctypes.CDLL(lib_path)                       # loads without a GPU; no compute call is made
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/fsm_b200.h but not exported"
    from torchfsm_b200 import _cabi
    assert set(_cabi.EXPORTS) == declared
    lib.fsm_abi_version.restype = ctypes.c_int
    lib.fsm_backend.restype = ctypes.c_int
    assert lib.fsm_abi_version() == 1 and lib.fsm_backend() == 0


def test_missing_cuda_library_is_a_loud_error(monkeypatch, tmp_path):
    from torchfsm_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "DEFAULT_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/torch fallback"):
        _cabi.lib()


def test_cpu_tensors_are_rejected_by_the_cuda_build():
    import torch
    import torchfsm_b200 as fsm
    from torchfsm_b200 import _cabi
    lib_path = os.path.join(PKG, "libfsm_b200.so")
    if not os.path.exists(lib_path):
        pytest.skip("libfsm_b200.so not built")
    _cabi.use_library(lib_path)
    try:
        u = torch.zeros(1, 1, 32, 32)
        with pytest.raises(RuntimeError, match="CUDA devices only"):
            (0.1 * fsm.Laplacian()).integrate(u, mesh=[(0, 1, 32), (0, 1, 32)], dt=0.1, step=1)
    finally:
        _cabi._lib = None


def test_multi_gpu_options_are_validated_before_any_work():
    """Host-side checks of the sharding API: bad exchange names, KS ensembles with an integrator whose zero-mode
    weights are not known, and the direct-exchange registration on a plan without slab decomposition."""
    import torch
    import torchfsm_b200 as fsm
    from product_util import build_emulator
    from torchfsm_b200 import _cabi
    _cabi.use_library(build_emulator())
    op = fsm.pde.NavierStokes(Re=100)
    with pytest.raises(ValueError):
        op.set_slab_decomposition(rank=0, nranks=2, exchange="carrier-pigeon")
    mesh = fsm.MeshGrid([(0, 20, 16)] * 2, device="cpu", dtype=torch.float64)
    ks = fsm.pde.KuramotoSivashinskyHighDim()
    ks.set_integrator(fsm.RKIntegrator.RK4)
    ks.set_ensemble_group(object())                      # any non-None group: the check happens before the collective
    u0 = torch.randn(2, 1, 16, 16, dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        ks.integrate(u0, mesh=mesh, dt=1e-3, step=1)
    ks2 = fsm.pde.KuramotoSivashinskyHighDim()
    ks2.integrate(u0, mesh=mesh, dt=1e-2, step=1)
    st = ks2._state_dict["integrator"]
    ptrs = (ctypes.c_void_p * 2)(1, 2)
    assert st._lib.fsm_slab_peers(st._plan, 1, ptrs, 2) != 0           # no slab decomposition on this plan
    assert b"slab" in st._lib.fsm_last_error()
    # a complex linear symbol on a 2-D grid runs as a pair of half spectra (unrolled.py) -- never silently dropped; with
    # a nonlinear term it needs the Nyquist planes dealiased away, and it does not run multi-GPU
    adv = -1.0 * fsm.SpatialDerivative(0, 1) + 0.01 * fsm.Laplacian()
    moved = adv.integrate(u0, mesh=mesh, dt=1e-2, step=1)
    still = (0.01 * fsm.Laplacian()).integrate(u0, mesh=mesh, dt=1e-2, step=1)
    assert float((moved - still).abs().max()) > 1e-3
    full = fsm.pde.KuramotoSivashinskyHighDim() + 0.1 * fsm.SpatialDerivative(0, 3)
    full.set_de_aliasing_rate(1.0)
    with pytest.raises(NotImplementedError):
        full.integrate(u0, mesh=mesh, dt=1e-3, step=1)
    sharded = fsm.pde.KuramotoSivashinskyHighDim() + 0.1 * fsm.SpatialDerivative(0, 3)
    sharded.set_ensemble_group(object())
    with pytest.raises(NotImplementedError):
        sharded.integrate(u0, mesh=mesh, dt=1e-3, step=1)


def test_no_tuning_hooks_in_the_product_path():
    """The CUDA build reads no environment variable (the only getenv sits behind FSM_EMU, a seam of the CPU test-suite)
    and the Python layer knows exactly one: FSM_B200_LIB, the library path used by tools/ to compare kernel builds.
    What bench.py runs is therefore what the defaults in the sources say."""
    import re
    csrc = os.path.join(ROOT, "torchfsm_b200", "csrc")
    for name in os.listdir(csrc):
        if not name.endswith((".cu", ".cuh", ".h")):
            continue
        lines = open(os.path.join(csrc, name)).read().split("\n")
        for i, line in enumerate(lines):
            if "getenv" in line:
                assert "#ifdef FSM_EMU" in lines[i - 1], f"{name}:{i + 1} reads the environment in the CUDA build"
    pkg = os.path.join(ROOT, "torchfsm_b200")
    seen = set()
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            seen |= set(re.findall(r"environ(?:\.get)?\(?\[?[\"']([A-Z0-9_]+)[\"']", open(os.path.join(pkg, name)).read()))
    assert seen == {"FSM_B200_LIB"}, seen
    if "FSM_B200_LIB" not in os.environ:
        from torchfsm_b200 import _cabi
        assert _cabi.DEFAULT_LIB == os.path.join(pkg, "libfsm_b200.so")
