#!/usr/bin/env python
"""Generate the golden input/output vectors that pin ``oracle/`` and the table builders.

Runs ONLY in the authoring container: it imports the UNMODIFIED reference from
``/root/reference`` (read-only) on CPU, executes small cases of the hot path
(``Operator.integrate`` -> integrator.step -> nonlinear evaluation -> FFTs) and stores
inputs, coefficient tables and outputs as ``tests/golden/<case>_<dtype>.npz``.
The reference cannot travel to the GPU box, the fixtures do.

    python tests/golden/make_golden.py            # regenerates every fixture

Each fixture stores: ``spec`` (json: mesh, terms, integrator, dt, steps, batch, channels),
``u0``; reference ``linear_coef``; every integrator table (``tab_*``); ``n0_hat`` = one
nonlinear evaluation of fft(u0); ``u1_hat`` / ``u1`` = state after ONE step; ``uT_hat`` /
``uT`` = state after all steps. Source arrays of explicit-source terms are stored as
``src_<i>``.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
import torchfsm  # noqa: E402
from torchfsm.mesh import MeshGrid, FourierMesh  # noqa: E402
from torchfsm.operator import (Operator, Laplacian, Biharmonic, Convection, KSConvection,  # noqa: E402
                               VorticityConvection, NSPressureConvection, ImplicitSource, ExplicitSource,
                               SpatialDerivative)
from torchfsm.integrator import ETDRKIntegrator, SETDRKIntegrator, RKIntegrator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)

INTEGRATORS = {
    "ETDRK0": ETDRKIntegrator.ETDRK0, "ETDRK1": ETDRKIntegrator.ETDRK1, "ETDRK2": ETDRKIntegrator.ETDRK2,
    "SETDRK1": SETDRKIntegrator.SETDRK1, "SETDRK2": SETDRKIntegrator.SETDRK2,
    "SETDRK3": SETDRKIntegrator.SETDRK3, "SETDRK4": SETDRKIntegrator.SETDRK4,
    "RK4": RKIntegrator.RK4,
}


def build_reference_operator(terms, sources):
    """terms: [(kind, coef, params)] -> reference Operator built from its public classes."""
    op = None
    for i, (kind, coef, params) in enumerate(terms):
        if kind == "laplacian":
            t = Laplacian()
        elif kind == "biharmonic":
            t = Biharmonic()
        elif kind == "spatial_derivative":
            t = SpatialDerivative(params["dim_index"], params["order"])
        elif kind == "implicit_unit_source":
            t = ImplicitSource()
        elif kind == "convection":
            t = Convection()
        elif kind == "ks_convection":
            t = KSConvection(params.get("remove_mean", True))
        elif kind == "vorticity_convection":
            t = VorticityConvection()
        elif kind == "ns_pressure_convection":
            t = NSPressureConvection()
        elif kind == "explicit_source":
            t = ExplicitSource(sources[i])
        else:
            raise ValueError(kind)
        t = coef * t
        op = t if op is None else op + t
    return op


def initial_condition(name, mesh_info, batch, channels, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    mg = MeshGrid(mesh_info, dtype=dtype)
    shape = [batch, channels] + [m[2] for m in mesh_info]
    if name == "burgers1d_readme":                       # README.md:45-48
        x = mg.bc_mesh_grid()
        return torch.sin(2 * torch.pi * x) + 0.5
    if name == "smooth_noise":
        # band-limited random field: random low modes, real, O(1) magnitude
        u = torch.randn(*shape, generator=g, dtype=dtype)
        fm = FourierMesh(mesh_info, dtype=dtype)
        u_hat = fm.fft(u) * fm.low_pass_filter(0.4)
        u = fm.ifft(u_hat).real
        return u / u.abs().amax(dim=tuple(range(1, u.ndim)), keepdim=True)
    if name == "white_noise":
        return torch.randn(*shape, generator=g, dtype=dtype)
    if name == "taylor_green3d":                         # ns_velocity.ipynb cell 2
        x, y, z = mg.bc_mesh_grid()
        u = torch.cat([torch.sin(x) * torch.cos(y) * torch.cos(z),
                       -torch.cos(x) * torch.sin(y) * torch.cos(z),
                       torch.zeros_like(x)], dim=1)
        pert = torch.randn(*shape, generator=g, dtype=dtype) * 0.05
        fm = FourierMesh(mesh_info, dtype=dtype)
        pert = fm.ifft(fm.fft(pert) * fm.low_pass_filter(0.5)).real
        return u.repeat(batch, 1, 1, 1, 1) + pert
    raise ValueError(name)


def make_sources(terms, mesh_info, dtype):
    """Explicit-source arrays: params['source'] names a recipe."""
    srcs = {}
    mg = MeshGrid(mesh_info, dtype=dtype)
    for i, (kind, coef, params) in enumerate(terms):
        if kind != "explicit_source":
            continue
        recipe = params["source"]
        if recipe == "kolm_y":                           # field.py:145-148 with x := y grid, k=4
            grids = mg.bc_mesh_grid()
            y = grids[1]
            srcs[i] = 4.0 * torch.cos(4.0 * 1.0 * y)
        else:
            raise ValueError(recipe)
    return srcs


TWO_PI = 2 * np.pi
CASES = [
    # name, mesh_info, batch, channels, terms, integrator, dt, steps, ic
    dict(name="c1_burgers1d_128", mesh=[(0, 1, 128)], B=1, C=1,
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})],
         integrator="auto", dt=0.01, steps=200, ic="burgers1d_readme"),
    dict(name="burgers1d_64_etdrk1", mesh=[(0, 1, 64)], B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})],
         integrator="ETDRK1", dt=0.005, steps=4, ic="smooth_noise"),
    dict(name="burgers1d_64_etdrk2", mesh=[(0, 1, 64)], B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})],
         integrator="ETDRK2", dt=0.005, steps=4, ic="smooth_noise"),
    dict(name="burgers1d_64_setdrk1", mesh=[(0, 1, 64)], B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})],
         integrator="SETDRK1", dt=0.005, steps=4, ic="smooth_noise"),
    dict(name="burgers1d_64_setdrk2", mesh=[(0, 1, 64)], B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})],
         integrator="SETDRK2", dt=0.005, steps=4, ic="smooth_noise"),
    dict(name="burgers1d_64_setdrk3", mesh=[(0, 1, 64)], B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})],
         integrator="SETDRK3", dt=0.005, steps=4, ic="smooth_noise"),
    dict(name="burgers1d_64_rk4", mesh=[(0, 1, 64)], B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})],
         integrator="RK4", dt=0.0005, steps=4, ic="smooth_noise"),
    dict(name="diffusion2d_32_etdrk0", mesh=[(0, 1, 32), (0, 2, 32)], B=2, C=2,
         terms=[("laplacian", 0.05, {})],
         integrator="auto", dt=0.1, steps=3, ic="white_noise"),
    dict(name="c2_ks2d_32", mesh=[(0, 30, 32), (0, 30, 32)], B=3, C=1,
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {"remove_mean": True})],
         integrator="SETDRK4", dt=0.1, steps=5, ic="smooth_noise"),
    dict(name="c3_ns2d_32_etdrk2", mesh=[(0, TWO_PI, 32), (0, TWO_PI, 32)], B=2, C=1,
         terms=[("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {}),
                ("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": "kolm_y"})],
         integrator="ETDRK2", dt=0.01, steps=5, ic="smooth_noise"),
    dict(name="c3_ns2d_64x32_etdrk2", mesh=[(0, TWO_PI, 64), (0, TWO_PI, 32)], B=2, C=1,
         terms=[("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {}),
                ("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": "kolm_y"})],
         integrator="ETDRK2", dt=0.01, steps=5, ic="smooth_noise"),
    dict(name="ns2d_32_setdrk4", mesh=[(0, TWO_PI, 32), (0, TWO_PI, 32)], B=2, C=1,
         terms=[("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {})],
         integrator="SETDRK4", dt=0.01, steps=5, ic="smooth_noise"),
    dict(name="ns2d_32_rk4", mesh=[(0, TWO_PI, 32), (0, TWO_PI, 32)], B=2, C=1,
         terms=[("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {})],
         integrator="RK4", dt=0.002, steps=5, ic="smooth_noise"),
    dict(name="burgers2d_32x64_etdrk2", mesh=[(0, 1, 32), (0, 2, 64)], B=2, C=2,
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})],
         integrator="ETDRK2", dt=0.002, steps=4, ic="smooth_noise"),
    dict(name="c4_burgers3d_16", mesh=[(0, 1, 16)] * 3, B=2, C=3,
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})],
         integrator="SETDRK4", dt=0.002, steps=3, ic="smooth_noise"),
    dict(name="burgers3d_8x16x32_rk4", mesh=[(0, 1, 8), (0, 2, 16), (0, 1, 32)], B=1, C=3,
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})],
         integrator="RK4", dt=0.0005, steps=2, ic="smooth_noise"),
    dict(name="c5_ns3d_16_setdrk4", mesh=[(0, TWO_PI, 16)] * 3, B=1, C=3,
         terms=[("ns_pressure_convection", 1, {}), ("laplacian", 1 / 100, {})],
         integrator="SETDRK4", dt=0.0025, steps=3, ic="taylor_green3d"),
    dict(name="c5_ns3d_32x16x8_etdrk2", mesh=[(0, TWO_PI, 32), (0, TWO_PI, 16), (0, TWO_PI, 8)], B=2, C=3,
         terms=[("ns_pressure_convection", 1, {}), ("laplacian", 1 / 100, {})],
         integrator="ETDRK2", dt=0.0025, steps=3, ic="taylor_green3d"),
    dict(name="ks3d_16", mesh=[(0, 20, 16)] * 3, B=2, C=1,
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {"remove_mean": True})],
         integrator="SETDRK4", dt=0.05, steps=3, ic="smooth_noise"),
    # 1-D Kuramoto-Sivashinsky preset (pde.py:27-39) and odd-order linear terms: complex exp(L dt) (KdV, pde.py:51-64)
    dict(name="ks1d_64", mesh=[(0, 32, 64)], B=2, C=1,
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("convection", -1, {})],
         integrator="auto", dt=0.05, steps=4, ic="smooth_noise"),
    dict(name="kdv1d_64_setdrk4", mesh=[(0, 20, 64)], B=2, C=1,
         terms=[("spatial_derivative", -1, {"dim_index": 0, "order": 3}), ("convection", 6.0, {})],
         integrator="auto", dt=0.001, steps=4, ic="smooth_noise"),
    dict(name="kdv1d_64_etdrk2", mesh=[(0, 20, 64)], B=2, C=1,
         terms=[("spatial_derivative", -1, {"dim_index": 0, "order": 3}), ("convection", 6.0, {})],
         integrator="ETDRK2", dt=0.001, steps=4, ic="smooth_noise"),
    dict(name="kdv1d_64_rk4", mesh=[(0, 20, 64)], B=2, C=1,
         terms=[("spatial_derivative", -1, {"dim_index": 0, "order": 3}), ("convection", 6.0, {})],
         integrator="RK4", dt=0.0001, steps=4, ic="smooth_noise"),
    dict(name="advdiff1d_32_etdrk0", mesh=[(0, 1, 32)], B=2, C=1,
         terms=[("spatial_derivative", -0.7, {"dim_index": 0, "order": 1}), ("laplacian", 0.01, {})],
         integrator="ETDRK0", dt=0.01, steps=5, ic="smooth_noise"),
]


def run_case(case, dtype):
    mesh_info = [tuple(m) for m in case["mesh"]]
    srcs = make_sources(case["terms"], mesh_info, dtype)
    op = build_reference_operator(case["terms"], srcs)
    if case["integrator"] != "auto":
        op.set_integrator(INTEGRATORS[case["integrator"]])
    u0 = initial_condition(case["ic"], mesh_info, case["B"], case["C"], dtype, seed=1234)
    mesh = MeshGrid(mesh_info, dtype=dtype)
    dt, steps = case["dt"], case["steps"]

    out = {"u0": u0.numpy()}
    # one step (registers the mesh, builds the integrator)
    u1_hat = op.integrate(u0.clone(), mesh=mesh, dt=dt, step=1, return_in_fourier=True)
    sd = op._state_dict
    f_mesh = sd["f_mesh"]
    u0_hat = f_mesh.fft(u0)
    out["u1_hat"] = u1_hat.numpy()
    out["u1"] = f_mesh.ifft(u1_hat).real.numpy()
    if sd["linear_coef"] is not None:
        out["linear_coef"] = sd["linear_coef"].numpy()
    if sd["nonlinear_func"] is not None:
        out["n0_hat"] = sd["nonlinear_func"](u0_hat.clone()).numpy()
    integ = sd["integrator"]
    for attr in ("_exp_term", "_half_exp_term", "_coef_1", "_coef_2", "_coef_3", "_coef_4", "_coef_5", "_coef_6"):
        if hasattr(integ, attr):
            out["tab" + attr] = getattr(integ, attr).numpy()
    # all steps, integrator reused (mesh omitted => no rebuild)
    u_hat = u0_hat
    for _ in range(steps):
        u_hat = integ.forward(u_hat, dt)
    out["uT_hat"] = u_hat.numpy()
    out["uT"] = f_mesh.ifft(u_hat).real.numpy()
    # one right-hand-side evaluation through Operator.__call__
    out["rhs0"] = op(u0.clone()).numpy()
    for i, s in srcs.items():
        out[f"src_{i}"] = s.numpy()
    spec = dict(case)
    spec["dtype"] = str(dtype).replace("torch.", "")
    spec["torch"] = torch.__version__
    spec["reference"] = "qiauil/torchfsm v" + getattr(torchfsm, "__version__", "0.0.4")
    out["spec"] = np.array(json.dumps(spec))
    return out


def main():
    only = sys.argv[1:]
    for case in CASES:
        if only and case["name"] not in only:
            continue
        for dtype in (torch.float32, torch.float64):
            out = run_case(case, dtype)
            tag = "f32" if dtype == torch.float32 else "f64"
            path = os.path.join(HERE, f"{case['name']}_{tag}.npz")
            np.savez_compressed(path, **out)
            fin = np.isfinite(out["uT"]).all()
            print(f"{case['name']:32s} {tag} finite={fin} |uT|max={np.abs(out['uT']).max():.4g} "
                  f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
