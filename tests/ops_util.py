"""Case table and operator builder shared by tests/golden/make_golden_ops.py (which runs the cases on the UNMODIFIED
reference) and the product tests (emulator on CPU, CUDA library on the GPU). ``build_operator`` takes the module that
provides the operator classes -- ``torchfsm.operator`` there, ``torchfsm_b200`` here: same names, same arguments."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_OPS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ops")
TWO_PI = 2 * np.pi


def _grids(mesh_info, dtype, device="cpu"):
    """Broadcastable coordinate grids (1, 1, n0, 1, ..), the reference's MeshGrid.bc_mesh_grid (mesh.py:56-62, 120-160)."""
    nd = len(mesh_info)
    out = []
    for i, (a, b, n) in enumerate(mesh_info):
        ax = (b - a) * torch.arange(n, dtype=dtype, device=device) / n
        shape = [1, 1] + [1] * nd
        shape[2 + i] = n
        out.append(ax.reshape(shape))
    return out


def source_field(recipe, mesh_info, dtype, device="cpu"):
    g = _grids(mesh_info, dtype, device)
    full = [1, 1] + [m[2] for m in mesh_info]
    if recipe == "kolm_y":                     # field.py:145-148 with x := y grid, k = 4
        return (4.0 * torch.cos(4.0 * g[1])).expand(full).contiguous()
    if recipe == "force3d":                    # a smooth 3-component body force
        x, y, z = g
        return torch.cat([0.3 * torch.sin(y) * torch.cos(2 * z) + 0 * x, 0.2 * torch.cos(x + z) + 0 * y,
                          0.1 * torch.sin(2 * x) * torch.sin(y) + 0 * z], dim=1).contiguous()
    if recipe == "force2d":
        x, y = g
        return torch.cat([0.3 * torch.sin(2 * y) + 0 * x, 0.2 * torch.cos(x + y)], dim=1).contiguous()
    if recipe == "heat3d":
        x, y, z = g
        return (torch.sin(x) * torch.cos(2 * y) * torch.cos(z)).contiguous()
    if recipe == "heat2d":
        x, y = g
        return (torch.sin(x) * torch.cos(2 * y)).contiguous()
    raise ValueError(recipe)


FUNCS = {
    "allen_cahn": lambda u: u - u ** 3,
    "sin": lambda u: torch.sin(u),
}


def case_sources(case, dtype, device="cpu"):
    mesh_info = [tuple(m) for m in case["mesh"]]
    return lambda recipe: source_field(recipe, mesh_info, dtype, device)


def smooth_field(case, dtype):
    """Band-limited random field of O(1) magnitude (torch only; stored in the fixture as u0)."""
    g = torch.Generator().manual_seed(case.get("seed", 4321))
    shape = [case["B"], case["C"]] + [m[2] for m in case["mesh"]]
    u = torch.randn(*shape, generator=g, dtype=torch.float64)
    dims = tuple(range(2, len(shape)))
    u_hat = torch.fft.fftn(u, dim=dims)
    for d in dims:
        n = shape[d]
        f = torch.fft.fftfreq(n, 1.0 / n).abs()
        view = [1] * len(shape)
        view[d] = n
        u_hat = u_hat * (f <= max(2, int(0.3 * n / 2))).to(u_hat.dtype).reshape(view)
    u = torch.fft.ifftn(u_hat, dim=dims).real
    u = u / u.abs().amax(dim=tuple(range(1, u.ndim)), keepdim=True)
    if case.get("rough"):                      # white noise on top: every mode populated, the Nyquist planes included
        u = u + case["rough"] * torch.randn(*shape, generator=g, dtype=torch.float64)
    return u.to(dtype)


def build_operator(ops, terms, sources, dtype, device="cpu"):
    """terms = [(kind, coef, params)]: coef a number or a list (one value per sample -> tensor-valued coefficient of
    shape (B, 1, 1, ..)); params may hold 'source' / 'func' recipes and 'force' (a nested term list)."""
    op = None
    for kind, coef, params in terms:
        params = dict(params)
        force = None
        if params.get("force") is not None:
            force = build_operator(ops, params["force"], sources, dtype, device)
        if kind == "laplacian":
            t = ops.Laplacian()
        elif kind == "biharmonic":
            t = ops.Biharmonic()
        elif kind == "spatial_derivative":
            t = ops.SpatialDerivative(params["dim_index"], params["order"])
        elif kind == "implicit_unit_source":
            t = ops.ImplicitSource()
        elif kind == "implicit_func_source":
            t = ops.ImplicitSource(FUNCS[params["func"]], params.get("non_linear", True))
        elif kind == "explicit_source":
            t = ops.ExplicitSource(sources(params["source"]).to(device))
        elif kind == "convection":
            t = ops.Convection()
        elif kind == "conservative_convection":
            t = ops.ConservativeConvection()
        elif kind == "ks_convection":
            t = ops.KSConvection(params.get("remove_mean", True))
        elif kind == "vorticity_convection":
            t = ops.VorticityConvection()
        elif kind == "ns_pressure_convection":
            t = ops.NSPressureConvection(force)
        elif kind == "grad":
            t = ops.Grad()
        elif kind == "div":
            t = ops.Div()
        elif kind == "curl":
            t = ops.Curl()
        elif kind == "vorticity2velocity":
            t = ops.Vorticity2Velocity()
        elif kind == "vorticity2pressure":
            t = ops.Vorticity2Pressure(force)
        elif kind == "velocity2pressure":
            t = ops.Velocity2Pressure(force)
        else:
            raise ValueError(kind)
        if isinstance(coef, (list, tuple)):
            nd = params.pop("_ndim")
            c = torch.tensor(coef, dtype=dtype, device=device)
            coef = c.reshape([len(coef), 1] + [1] * nd)
        t = coef * t
        op = t if op is None else op + t
    return op


def _m(n, *lengths):
    return [(0, length, k) for length, k in zip(lengths, n)]


KOLM = [("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": "kolm_y"})]
OPS_CASES = [
    # ---- one evaluation: channel-changing cores (derivative_example.ipynb) and the NS diagnostics (ns_vorticity.ipynb)
    dict(name="grad2d", mode="call", mesh=_m((32, 16), TWO_PI, 3.0), B=2, C=1, terms=[("grad", 1, {})]),
    dict(name="grad3d", mode="call", mesh=_m((8, 16, 8), 1.0, 2.0, TWO_PI), B=2, C=1, terms=[("grad", -0.5, {})]),
    dict(name="grad1d", mode="call", mesh=_m((32,), 2.0), B=2, C=1, terms=[("grad", 1, {})]),
    dict(name="div2d", mode="call", mesh=_m((16, 32), TWO_PI, 3.0), B=2, C=2, terms=[("div", 1, {})]),
    dict(name="div3d", mode="call", mesh=_m((16, 8, 8), 1.0, 2.0, TWO_PI), B=2, C=3, terms=[("div", 2.0, {})]),
    dict(name="curl2d", mode="call", mesh=_m((16, 32), TWO_PI, 3.0), B=2, C=2, terms=[("curl", 1, {})]),
    dict(name="curl3d", mode="call", mesh=_m((8, 16, 16), 1.0, 2.0, TWO_PI), B=2, C=3, terms=[("curl", 1, {})]),
    dict(name="lap_plus_d3_call2d", mode="call", mesh=_m((16, 16), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("laplacian", 0.5, {}), ("spatial_derivative", 0.1, {"dim_index": 1, "order": 3}),
                ("biharmonic", -0.01, {}), ("implicit_unit_source", 2.0, {})]),
    dict(name="v2v_2d", mode="call", mesh=_m((32, 16), TWO_PI, TWO_PI), B=2, C=1, terms=[("vorticity2velocity", 1, {})]),
    dict(name="vel2p_2d", mode="call", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=2, terms=[("velocity2pressure", 1, {})]),
    dict(name="vel2p_3d", mode="call", mesh=_m((16, 16, 8), TWO_PI, TWO_PI, TWO_PI), B=1, C=3,
         terms=[("velocity2pressure", 1, {})]),
    dict(name="vel2p_2d_force", mode="call", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=2,
         terms=[("velocity2pressure", 1, {"force": [("explicit_source", 1, {"source": "force2d"})]})]),
    dict(name="vor2p_2d", mode="call", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=1, terms=[("vorticity2pressure", 1, {})]),
    dict(name="vor2p_2d_kolm", mode="call", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("vorticity2pressure", 1, {"force": KOLM})]),
    dict(name="conscon2d_call", mode="call", mesh=_m((32, 16), TWO_PI, 3.0), B=2, C=2,
         terms=[("conservative_convection", 1, {})]),
    dict(name="run_operators2d", mode="run_operators", mesh=_m((16, 32), TWO_PI, 3.0), B=2, C=2,
         operators=[[("div", 1, {})], [("curl", 1, {})], [("laplacian", 0.1, {}), ("convection", -1, {})]]),
    dict(name="solve_third_derivative2d", mode="solve", mesh=_m((16, 32), TWO_PI, 3.0), B=2, C=1, rough=0.3,
         terms=[("spatial_derivative", 0.7, {"dim_index": 1, "order": 3})]),
    dict(name="solve_poisson2d", mode="solve", mesh=_m((32, 16), TWO_PI, 3.0), B=2, C=1, terms=[("laplacian", 1, {})]),
    # ---- time stepping
    dict(name="ns3d_force_setdrk4", mode="integrate", mesh=_m((16, 16, 16), TWO_PI, TWO_PI, TWO_PI), B=1, C=3,
         terms=[("ns_pressure_convection", 1, {"force": [("explicit_source", 1, {"source": "force3d"})]}),
                ("laplacian", 1 / 100, {})], integrator="SETDRK4", dt=0.0025, steps=3),
    dict(name="ns3d_force_etdrk2", mode="integrate", mesh=_m((16, 8, 16), TWO_PI, TWO_PI, TWO_PI), B=2, C=3,
         terms=[("ns_pressure_convection", 1, {"force": [("explicit_source", 2.0, {"source": "force3d"})]}),
                ("laplacian", 1 / 100, {})], integrator="ETDRK2", dt=0.0025, steps=3),
    dict(name="ns3d_force_etdrk1", mode="integrate", mesh=_m((16, 8, 16), TWO_PI, TWO_PI, TWO_PI), B=1, C=3,
         terms=[("ns_pressure_convection", 1, {"force": [("explicit_source", 1, {"source": "force3d"})]}),
                ("laplacian", 1 / 100, {})], integrator="ETDRK1", dt=0.0025, steps=3),
    dict(name="ns2d_velocity_setdrk4", mode="integrate", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=2,
         terms=[("ns_pressure_convection", 1, {}), ("laplacian", 1 / 100, {})], integrator="auto", dt=0.005, steps=3),
    dict(name="ns2d_velocity_force_etdrk2", mode="integrate", mesh=_m((32, 16), TWO_PI, TWO_PI), B=2, C=2,
         terms=[("ns_pressure_convection", 1, {"force": [("explicit_source", 1, {"source": "force2d"})]}),
                ("laplacian", 1 / 100, {})], integrator="ETDRK2", dt=0.005, steps=3),
    # state-dependent forces: drag + body force (the 3-D analogue of kolm_force), and a force with its own derivative.
    # fp32 only: the half-spectrum state reproduces the reference to ~1e-8 here (see DynamicForceStepper / DESIGN.md)
    dict(name="ns3d_drag_force_setdrk4", mode="integrate", dtypes=["float32"], mesh=_m((16, 16, 8), TWO_PI, TWO_PI, TWO_PI), B=2, C=3,
         terms=[("ns_pressure_convection", 1, {"force": [("implicit_unit_source", -0.1, {}),
                                                        ("explicit_source", 1, {"source": "force3d"})]}),
                ("laplacian", 1 / 100, {})], integrator="auto", dt=0.0025, steps=3),
    dict(name="ns2d_velocity_drag_force_etdrk2", mode="integrate", dtypes=["float32"], mesh=_m((32, 16), TWO_PI, TWO_PI), B=2, C=2,
         terms=[("ns_pressure_convection", 0.5, {"force": [("implicit_unit_source", -0.2, {}), ("laplacian", 0.01, {}),
                                                          ("explicit_source", 1, {"source": "force2d"})]}),
                ("laplacian", 1 / 100, {})], integrator="ETDRK2", dt=0.005, steps=3),
    dict(name="ns3d_drag_force_etdrk1", mode="integrate", dtypes=["float32"], mesh=_m((8, 16, 16), TWO_PI, TWO_PI, TWO_PI), B=1, C=3,
         terms=[("ns_pressure_convection", 1, {"force": [("implicit_unit_source", -0.1, {})]}),
                ("laplacian", 1 / 100, {})], integrator="ETDRK1", dt=0.0025, steps=3),
    # the explicit Runge-Kutta family other than RK4 (integrator/_rk.py:82-255), non-adaptive
    dict(name="burgers1d_euler", mode="integrate", mesh=_m((64,), 1.0), B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})], integrator="Euler", dt=0.0002, steps=3),
    dict(name="burgers1d_midpoint", mode="integrate", mesh=_m((64,), 1.0), B=2, C=1,
         terms=[("laplacian", 0.02, {}), ("convection", -1, {})], integrator="Midpoint", dt=0.0002, steps=3),
    dict(name="kdv1d_ralston12", mode="integrate", mesh=_m((64,), 20.0), B=2, C=1,
         terms=[("spatial_derivative", -1, {"dim_index": 0, "order": 3}), ("convection", 6.0, {})], integrator="Ralston12",
         dt=0.00005, steps=3),
    dict(name="burgers2d_heun12", mode="integrate", mesh=_m((32, 16), 1.0, 1.0), B=2, C=2,
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})], integrator="Heun12", dt=0.0005, steps=3),
    dict(name="ks2d_bogacki_shampine23", mode="integrate", mesh=_m((32, 32), 30.0, 30.0), B=2, C=1,
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {})], integrator="BogackiShampine23",
         dt=0.002, steps=3),
    dict(name="ns2d_kolm_rk4_38rule", mode="integrate", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("vorticity_convection", -1, {}), ("laplacian", 0.01, {})] + KOLM, integrator="RK4_38Rule", dt=0.002, steps=3),
    dict(name="burgers3d_dorpi45", mode="integrate", mesh=_m((16, 8, 8), 1.0, 1.0, 1.0), B=1, C=3,
         terms=[("laplacian", 0.01, {}), ("convection", -1, {})], integrator="Dorpi45", dt=0.0005, steps=3),
    dict(name="ns3d_fehlberg45", mode="integrate", mesh=_m((16, 16, 8), TWO_PI, TWO_PI, TWO_PI), B=1, C=3,
         terms=[("ns_pressure_convection", 1, {}), ("laplacian", 1 / 100, {})], integrator="Fehlberg45", dt=0.002, steps=3),
    dict(name="conscon2d_cashkarp45", mode="integrate", mesh=_m((32, 32), 1.0, 1.0), B=2, C=2,
         terms=[("laplacian", 0.01, {}), ("conservative_convection", -1, {})], integrator="CashKarp45", dt=0.0005, steps=3),
    dict(name="burgers2d_batched_nu_etdrk2", mode="integrate", mesh=_m((32, 16), 1.0, 1.0), B=3, C=2,
         terms=[("laplacian", [0.01, 0.02, 0.05], {"_ndim": 2}), ("convection", -1, {})], integrator="ETDRK2", dt=0.002, steps=3),
    dict(name="burgers1d_batched_nu_setdrk4", mode="integrate", mesh=_m((64,), 1.0), B=3, C=1,
         terms=[("laplacian", [0.01, 0.02, 0.05], {"_ndim": 1}), ("convection", -1, {})], integrator="auto", dt=0.002, steps=3),
    dict(name="burgers2d_batched_convection_coef", mode="integrate", mesh=_m((32, 16), 1.0, 1.0), B=3, C=2,
         terms=[("laplacian", 0.01, {}), ("convection", [-1.0, -0.5, -2.0], {"_ndim": 2})], integrator="ETDRK2", dt=0.002, steps=3),
    dict(name="kdv1d_batched_convection_coef", mode="integrate", mesh=_m((64,), 20.0), B=2, C=1,
         terms=[("spatial_derivative", -1, {"dim_index": 0, "order": 3}), ("convection", [6.0, 3.0], {"_ndim": 1})],
         integrator="auto", dt=0.001, steps=3),
    dict(name="ns3d_batched_coef_setdrk4", mode="integrate", mesh=_m((16, 8, 16), TWO_PI, TWO_PI, TWO_PI), B=2, C=3,
         terms=[("ns_pressure_convection", [1.0, 0.5], {"_ndim": 3}), ("laplacian", 1 / 100, {})], integrator="auto",
         dt=0.0025, steps=3),
    dict(name="conscon2d_batched_coef_etdrk2", mode="integrate", mesh=_m((32, 16), 1.0, 1.0), B=2, C=2,
         terms=[("laplacian", 0.01, {}), ("conservative_convection", [-1.0, -0.5], {"_ndim": 2})], integrator="ETDRK2",
         dt=0.002, steps=3),
    # ---- complex linear symbol (odd-order terms) on 2-D/3-D grids: the reference's state leaves the Hermitian subspace on
    # the Nyquist planes; rough initial data so that those planes matter (torchfsm_b200/unrolled.py)
    dict(name="advdiff2d_complex_etdrk0", mode="integrate", mesh=_m((16, 32), 1.0, 1.0), B=2, C=1, rough=0.3,
         terms=[("laplacian", 0.01, {}), ("spatial_derivative", 0.7, {"dim_index": 0, "order": 1}),
                ("spatial_derivative", -0.3, {"dim_index": 1, "order": 1})], integrator="auto", dt=0.01, steps=4),
    dict(name="advdiff2d_complex_rk4", mode="integrate", mesh=_m((16, 16), 1.0, 1.0), B=2, C=1, rough=0.3,
         terms=[("laplacian", 0.01, {}), ("spatial_derivative", 0.7, {"dim_index": 0, "order": 1})],
         integrator="RK4", dt=0.0005, steps=3),
    dict(name="dispersion3d_source_setdrk4", mode="integrate", mesh=_m((8, 16, 8), TWO_PI, TWO_PI, TWO_PI), B=1, C=1, rough=0.3,
         terms=[("laplacian", 0.01, {}), ("spatial_derivative", 0.7, {"dim_index": 2, "order": 1}),
                ("spatial_derivative", -0.05, {"dim_index": 0, "order": 3}), ("explicit_source", 1, {"source": "heat3d"})],
         integrator="auto", dt=0.002, steps=3),
    dict(name="beta_plane2d_etdrk2", mode="integrate", mesh=_m((32, 16), TWO_PI, TWO_PI), B=2, C=1, rough=0.2,
         terms=[("laplacian", 0.01, {}), ("vorticity_convection", -1, {}),
                ("spatial_derivative", 0.5, {"dim_index": 0, "order": 1})] + KOLM, integrator="ETDRK2", dt=0.01, steps=3),
    dict(name="beta_plane2d_setdrk4", mode="integrate", mesh=_m((16, 16), TWO_PI, TWO_PI), B=2, C=1, rough=0.2,
         terms=[("laplacian", 0.01, {}), ("vorticity_convection", -1, {}),
                ("spatial_derivative", 0.5, {"dim_index": 0, "order": 1})], integrator="SETDRK4", dt=0.01, steps=3),
    dict(name="ks2d_dispersion_setdrk3", mode="integrate", mesh=_m((16, 16), 10.0, 10.0), B=2, C=1, rough=0.2,
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {}),
                ("spatial_derivative", 0.3, {"dim_index": 0, "order": 3})], integrator="SETDRK3", dt=0.02, steps=3),
    dict(name="allen_cahn2d_advection_setdrk2", mode="integrate", mesh=_m((16, 16), 1.0, 1.0), B=2, C=1, rough=0.2,
         terms=[("laplacian", 0.05, {}), ("implicit_func_source", 1, {"func": "allen_cahn"}),
                ("spatial_derivative", 0.4, {"dim_index": 1, "order": 1})], integrator="SETDRK2", dt=0.01, steps=3),
    dict(name="beta_plane2d_dorpi45", mode="integrate", mesh=_m((16, 16), TWO_PI, TWO_PI), B=2, C=1, rough=0.2,
         terms=[("laplacian", 0.01, {}), ("vorticity_convection", -1, {}),
                ("spatial_derivative", 0.5, {"dim_index": 0, "order": 1})], integrator="Dorpi45", dt=0.002, steps=3),
    dict(name="ks2d_batched_setdrk4", mode="integrate", mesh=_m((32, 32), 30.0, 30.0), B=2, C=1,
         terms=[("laplacian", [-1.0, -0.9], {"_ndim": 2}), ("biharmonic", -1, {}), ("ks_convection", -1, {})],
         integrator="SETDRK4", dt=0.05, steps=3),
    dict(name="diffusion3d_batched_etdrk0", mode="integrate", mesh=_m((8, 8, 16), 1.0, 1.0, 2.0), B=2, C=1,
         terms=[("laplacian", [0.01, 0.03], {"_ndim": 3})], integrator="auto", dt=0.1, steps=3),
    dict(name="conscon2d_setdrk4", mode="integrate", mesh=_m((32, 32), 1.0, 1.0), B=2, C=2,
         terms=[("laplacian", 0.01, {}), ("conservative_convection", -1, {})], integrator="auto", dt=0.002, steps=3),
    dict(name="conscon3d_etdrk2", mode="integrate", mesh=_m((16, 8, 8), 1.0, 1.0, 1.0), B=1, C=3,
         terms=[("laplacian", 0.01, {}), ("conservative_convection", -1, {})], integrator="ETDRK2", dt=0.002, steps=3),
    dict(name="conscon1d_rk4", mode="integrate", mesh=_m((64,), 1.0), B=2, C=1,
         terms=[("laplacian", 0.01, {}), ("conservative_convection", -0.5, {})], integrator="RK4", dt=0.0005, steps=3),
    dict(name="ksconv1d_setdrk4", mode="integrate", mesh=_m((64,), 32.0), B=3, C=1,
         terms=[("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {})], integrator="auto", dt=0.05, steps=3),
    dict(name="allen_cahn2d_etdrk2", mode="integrate", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("laplacian", 0.05, {}), ("implicit_func_source", 1, {"func": "allen_cahn"})],
         integrator="ETDRK2", dt=0.01, steps=3),
    dict(name="sine_source2d_linear_flag_setdrk4", mode="integrate", mesh=_m((16, 32), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("laplacian", 0.05, {}), ("implicit_func_source", 0.5, {"func": "sin", "non_linear": False})],
         integrator="auto", dt=0.01, steps=3),
    dict(name="heat2d_explicit_source", mode="integrate", mesh=_m((32, 16), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("laplacian", 0.05, {}), ("explicit_source", 1, {"source": "heat2d"})], integrator="auto", dt=0.05, steps=3),
    dict(name="ns2d_fourth_derivative_etdrk2", mode="integrate", mesh=_m((32, 32), TWO_PI, TWO_PI), B=2, C=1,
         terms=[("vorticity_convection", -1, {}), ("spatial_derivative", 0.01, {"dim_index": 0, "order": 2}),
                ("spatial_derivative", -0.001, {"dim_index": 1, "order": 4})], integrator="ETDRK2", dt=0.01, steps=3),
]


def ops_names(dtype_tag=None, mode=None):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_OPS_DIR, "*.npz")))
    if dtype_tag:
        names = [n for n in names if n.endswith("_" + dtype_tag)]
    if mode:
        by = {c["name"]: c["mode"] for c in OPS_CASES}
        names = [n for n in names if by.get(n[:-4]) in (mode if isinstance(mode, (tuple, list)) else (mode,))]
    return names


def load_ops(name):
    z = np.load(os.path.join(GOLDEN_OPS_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["spec"] = json.loads(str(g["spec"]))
    return g
