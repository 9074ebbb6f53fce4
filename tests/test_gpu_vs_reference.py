"""Product against the UNMODIFIED reference on the same GPU (the primary oracle is the imported
reference executed on the same inputs, same device, same dtype). The reference travels to the GPU box as baseline/_ref
(pip --target install, git-ignored); the tests skip where it is absent. Operators are built from the same term lists by
the same builder for both packages (tests/ops_util.py: identical class names), tables are built by each side with the
same torch expressions on the same device, so the comparison is apples to apples at sizes the fixtures do not reach."""
import os
import sys

import pytest
import torch

from ops_util import build_operator, case_sources, smooth_field, TWO_PI

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "torchfsm")):
            if cand not in sys.path:
                sys.path.append(cand)
            try:
                import torchfsm  # noqa: F401
                import torchfsm.operator as ref_ops
                return torchfsm, ref_ops
            except Exception:
                continue
    return None, None


def _m(n, *lengths):
    return [(0, length, k) for length, k in zip(lengths, n)]


CASES = [
    dict(name="burgers1d_1024_setdrk4", mesh=_m((1024,), 1.0), B=4, C=1, dt=1e-4, steps=3, integrator="SETDRK4",
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})]),
    dict(name="burgers2d_256x128_etdrk2", mesh=_m((256, 128), 1.0, 2.0), B=3, C=2, dt=1e-3, steps=3, integrator="ETDRK2",
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})]),
    dict(name="burgers3d_64_rk4", mesh=_m((64, 64, 64), 1.0, 1.0, 1.0), B=2, C=3, dt=2e-4, steps=2, integrator="RK4",
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})]),
    dict(name="ks2d_256_setdrk4", mesh=_m((256, 256), 60.0, 60.0), B=4, C=1, dt=0.25, steps=3, integrator="SETDRK4",
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {})]),
    dict(name="ks3d_64_setdrk3", mesh=_m((64, 32, 64), 40.0, 20.0, 40.0), B=2, C=1, dt=0.05, steps=3, integrator="SETDRK3",
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {})]),
    dict(name="ns2d_kolm_512_etdrk2", mesh=_m((512, 512), TWO_PI, TWO_PI), B=2, C=1, dt=0.01, steps=3, integrator="ETDRK2",
         terms=[("vorticity_convection", -1, {}), ("laplacian", 0.01, {}), ("implicit_unit_source", -0.1, {}),
                ("explicit_source", -1, {"source": "kolm_y"})]),
    dict(name="ns3d_64_setdrk4", mesh=_m((64, 64, 64), TWO_PI, TWO_PI, TWO_PI), B=1, C=3, dt=0.005, steps=3, integrator="SETDRK4",
         terms=[("ns_pressure_convection", 1, {}), ("laplacian", 1 / 400, {})]),
    dict(name="ns3d_force_64x32x64_setdrk4", mesh=_m((64, 32, 64), TWO_PI, TWO_PI, TWO_PI), B=2, C=3, dt=0.005, steps=3,
         integrator="SETDRK4",
         terms=[("ns_pressure_convection", 1, {"force": [("explicit_source", 1, {"source": "force3d"})]}),
                ("laplacian", 1 / 400, {})]),
    dict(name="ns2d_velocity_drag_256_etdrk2", mesh=_m((256, 256), TWO_PI, TWO_PI), B=2, C=2, dt=0.005, steps=3,
         integrator="ETDRK2", dtypes=["float32"],
         terms=[("ns_pressure_convection", 1, {"force": [("implicit_unit_source", -0.1, {}),
                                                        ("explicit_source", 1, {"source": "force2d"})]}),
                ("laplacian", 1 / 400, {})]),
    # complex linear symbol on 2-D/3-D grids (paired half spectra, torchfsm_b200/unrolled.py), rough data
    dict(name="beta_plane2d_256_etdrk2", mesh=_m((256, 256), TWO_PI, TWO_PI), B=2, C=1, dt=0.005, steps=3, integrator="ETDRK2",
         rough=0.05, terms=[("vorticity_convection", -1, {}), ("laplacian", 0.01, {}),
                            ("spatial_derivative", 0.5, {"dim_index": 0, "order": 1})]),
    dict(name="advection_dispersion3d_64_setdrk4", mesh=_m((64, 32, 64), TWO_PI, TWO_PI, TWO_PI), B=2, C=1, dt=0.002, steps=3,
         integrator="SETDRK4", rough=0.05,
         terms=[("laplacian", 0.01, {}), ("spatial_derivative", 0.7, {"dim_index": 2, "order": 1}),
                ("spatial_derivative", -0.01, {"dim_index": 0, "order": 3}), ("ks_convection", -1, {})]),
    dict(name="conscon2d_256_setdrk4", mesh=_m((256, 256), 1.0, 1.0), B=2, C=2, dt=1e-3, steps=3, integrator="SETDRK4",
         terms=[("laplacian", 0.01, {}), ("conservative_convection", -1, {})]),
    dict(name="allen_cahn3d_32_etdrk2", mesh=_m((32, 64, 32), TWO_PI, TWO_PI, TWO_PI), B=2, C=1, dt=0.01, steps=3,
         integrator="ETDRK2", terms=[("laplacian", 0.05, {}), ("implicit_func_source", 1, {"func": "allen_cahn"})]),
    dict(name="burgers2d_batched_nu_256_setdrk4", mesh=_m((256, 256), 1.0, 1.0), B=3, C=2, dt=1e-3, steps=3, integrator="SETDRK4",
         terms=[("laplacian", [0.01, 0.02, 0.05], {"_ndim": 2}), ("convection", -1, {})]),
]
CALLS = [
    dict(name="curl3d_64", mesh=_m((64, 64, 32), 1.0, 2.0, TWO_PI), B=2, C=3, terms=[("curl", 1, {})]),
    dict(name="grad2d_512", mesh=_m((512, 256), TWO_PI, 3.0), B=2, C=1, terms=[("grad", 1, {})]),
    dict(name="vor2p_kolm_256", mesh=_m((256, 256), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("vorticity2pressure", 1, {"force": [("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": "kolm_y"})]})]),
    dict(name="vel2p_3d_64", mesh=_m((64, 64, 64), TWO_PI, TWO_PI, TWO_PI), B=1, C=3, terms=[("velocity2pressure", 1, {})]),
]


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def _enum(pkg, name):
    for enum in (pkg.ETDRKIntegrator, pkg.SETDRKIntegrator, pkg.RKIntegrator):
        if name in enum.__members__:
            return enum[name]
    raise ValueError(name)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_steps_match_the_reference_on_this_gpu(case, dtype):
    torchfsm, ref_ops = _reference()
    if torchfsm is None:
        pytest.skip("reference not importable on this box (baseline/_ref did not travel)")
    if str(dtype).replace("torch.", "") not in case.get("dtypes", ["float32", "float64"]):
        pytest.skip("this case runs in fp32 only (state-dependent force)")
    import torchfsm.integrator as ref_int
    import torchfsm_b200 as fsm
    dev = "cuda"
    tol = 1e-5 if dtype == torch.float32 else 1e-11
    u0 = smooth_field(case, dtype).to(dev)
    srcs = case_sources(case, dtype, dev)
    ref = build_operator(ref_ops, case["terms"], srcs, dtype, dev)
    ref.set_integrator(_enum(ref_int, case["integrator"]))
    ours = build_operator(fsm, case["terms"], srcs, dtype, dev)
    ours.set_integrator(_enum(fsm, case["integrator"]))
    ref_mesh = torchfsm.mesh.MeshGrid([tuple(m) for m in case["mesh"]], device=dev, dtype=dtype)
    our_mesh = fsm.MeshGrid([tuple(m) for m in case["mesh"]], device=dev, dtype=dtype)
    want = ref.integrate(u0.clone(), mesh=ref_mesh, dt=case["dt"], step=1)
    got = ours.integrate(u0.clone(), mesh=our_mesh, dt=case["dt"], step=1)
    assert _rel(got, want) <= tol, "first step"
    want = ref.integrate(u0.clone(), dt=case["dt"], step=case["steps"])
    got = ours.integrate(u0.clone(), dt=case["dt"], step=case["steps"])
    assert torch.isfinite(want).all()
    assert _rel(got, want) <= tol * case["steps"], "all steps"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("case", CALLS, ids=[c["name"] for c in CALLS])
def test_evaluations_match_the_reference_on_this_gpu(case, dtype):
    torchfsm, ref_ops = _reference()
    if torchfsm is None:
        pytest.skip("reference not importable on this box (baseline/_ref did not travel)")
    import torchfsm_b200 as fsm
    dev = "cuda"
    tol = 1e-5 if dtype == torch.float32 else 1e-11
    u0 = smooth_field(case, dtype).to(dev)
    srcs = case_sources(case, dtype, dev)
    ref = build_operator(ref_ops, case["terms"], srcs, dtype, dev)
    ours = build_operator(fsm, case["terms"], srcs, dtype, dev)
    want = ref(u0.clone(), mesh=torchfsm.mesh.MeshGrid([tuple(m) for m in case["mesh"]], device=dev, dtype=dtype))
    got = ours(u0.clone(), mesh=fsm.MeshGrid([tuple(m) for m in case["mesh"]], device=dev, dtype=dtype))
    assert got.shape == want.shape and _rel(got, want) <= 10 * tol
