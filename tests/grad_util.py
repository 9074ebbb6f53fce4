"""Cases of the gradient checks (tests/golden_grad): nonlinear operators integrated for a few steps and evaluated once,
differentiated with respect to the initial field. The same builder makes the operator from the reference's namespace
(tests/golden/make_golden_grad.py, authoring container only) and from torchfsm_b200's."""
import json
import os
from types import SimpleNamespace

import numpy as np
import torch

GRAD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_grad")
TWO_PI = 6.283185307179586

GRAD_CASES = [
    dict(name="burgers1d_setdrk4", op="burgers", mesh=[(0, 1, 32)], C=1, integrator="SETDRK4", dt=0.01),
    dict(name="burgers2d_etdrk2", op="burgers", mesh=[(0, 1, 16), (0, 1, 32)], C=2, integrator="ETDRK2", dt=0.005),
    dict(name="burgers2d_batched_coef_setdrk2", op="burgers_batched", mesh=[(0, 1, 16), (0, 1, 16)], C=2,
         integrator="SETDRK2", dt=0.005),
    dict(name="burgers3d_setdrk3", op="burgers", mesh=[(0, 1, 8), (0, 1, 16), (0, 1, 8)], C=3, integrator="SETDRK3", dt=0.005),
    dict(name="ks2d_setdrk4", op="ks", mesh=[(0, 10, 16), (0, 10, 16)], C=1, integrator="SETDRK4", dt=0.05),
    # op(u) applies k^4 up to k = 10 to the fp32 rounding noise of the unexcited modes (on both sides): loose in fp32
    dict(name="ks1d_etdrk1", op="ks", mesh=[(0, 20, 64)], C=1, integrator="ETDRK1", dt=0.05, tol32={"z": 2e-2}),
    dict(name="kdv1d_etdrk2", op="kdv", mesh=[(0, TWO_PI, 64)], C=1, integrator="ETDRK2", dt=0.001, tol32={"z": 2e-3}),
    dict(name="ns2d_kolmogorov_etdrk2", op="ns2d_forced", mesh=[(0, TWO_PI, 32), (0, TWO_PI, 16)], C=1,
         integrator="ETDRK2", dt=0.01),
    dict(name="ns2d_setdrk3", op="ns2d", mesh=[(0, TWO_PI, 16), (0, TWO_PI, 32)], C=1, integrator="SETDRK3", dt=0.01),
    dict(name="ns3d_setdrk4", op="ns3d", mesh=[(0, 1, 8), (0, 1, 16), (0, 1, 8)], C=3, integrator="SETDRK4", dt=0.005),
    dict(name="ns2d_velocity_setdrk1", op="ns3d", mesh=[(0, 1, 16), (0, 1, 16)], C=2, integrator="SETDRK1", dt=0.005),
    dict(name="allen_cahn2d_etdrk2", op="allen_cahn", mesh=[(0, 1, 16), (0, 1, 16)], C=1, integrator="ETDRK2", dt=0.01),
    dict(name="conservative2d_setdrk2", op="conservative", mesh=[(0, 1, 16), (0, 1, 16)], C=2, integrator="SETDRK2", dt=0.005),
    # complex linear symbol on a 2-D grid (paired half spectra, torchfsm_b200/unrolled.py); rough data: Nyquist planes matter
    dict(name="beta_plane2d_setdrk4", op="beta_plane", mesh=[(0, TWO_PI, 16), (0, TWO_PI, 16)], C=1, integrator="SETDRK4",
         dt=0.01, rough=0.2),
    # gradients with respect to PARAMETERS (inverse problems): per-sample viscosity and convection coefficient, a source
    dict(name="burgers2d_learned_coefs_setdrk4", op="burgers_learned", mesh=[(0, 1, 16), (0, 1, 16)], C=2,
         integrator="SETDRK4", dt=0.005),
    dict(name="burgers2d_learned_coefs_rk4", op="burgers_learned", mesh=[(0, 1, 16), (0, 1, 16)], C=2,
         integrator="RK4", dt=0.001),
    dict(name="burgers2d_learned_scalars_setdrk4", op="burgers_learned_scalar", mesh=[(0, 1, 16), (0, 1, 16)], C=2,
         integrator="SETDRK4", dt=0.005),
    dict(name="heat2d_learned_source_setdrk2", op="heat_learned_source", mesh=[(0, 1, 16), (0, 1, 16)], C=1,
         integrator="SETDRK2", dt=0.01),
    dict(name="burgers2d_rk4", op="burgers", mesh=[(0, 1, 16), (0, 1, 16)], C=2, integrator="RK4", dt=0.001),
    dict(name="burgers2d_dorpi45", op="burgers", mesh=[(0, 1, 16), (0, 1, 16)], C=2, integrator="Dorpi45", dt=0.001),
]
STEPS = 3
BATCH = 2


def grad_names():
    return [c["name"] for c in GRAD_CASES]


def grad_case(name):
    return next(c for c in GRAD_CASES if c["name"] == name)


def namespace(kind):
    """Operators, integrator enums and MeshGrid of the reference ("reference") or of torchfsm_b200 ("b200")."""
    if kind == "reference":
        import torchfsm.operator as o
        import torchfsm.integrator as i
        from torchfsm.mesh import MeshGrid
        names = {k: getattr(o, k) for k in dir(o) if k[0].isupper()}
        names.update(MeshGrid=MeshGrid, ETDRKIntegrator=i.ETDRKIntegrator, SETDRKIntegrator=i.SETDRKIntegrator,
                     RKIntegrator=i.RKIntegrator)
        return SimpleNamespace(**names)
    import torchfsm_b200 as fsm
    return fsm


def build(ns, case, mesh, dtype, device="cpu"):
    """-> (operator, [parameter tensors that require grad])"""
    k = case["op"]
    params = []
    if k == "burgers":
        op = 0.01 * ns.Laplacian() - ns.Convection()
    elif k == "burgers_batched":
        c = torch.tensor([-1.0, -0.5], dtype=dtype, device=device).reshape(BATCH, 1, *([1] * len(case["mesh"])))
        op = 0.01 * ns.Laplacian() + c * ns.Convection()
    elif k == "burgers_learned":
        ones = [1] * len(case["mesh"])
        nu = torch.tensor([0.01, 0.03], dtype=dtype, device=device).reshape(BATCH, 1, *ones).requires_grad_(True)
        c = torch.tensor([-1.0, -0.5], dtype=dtype, device=device).reshape(BATCH, 1, *ones).requires_grad_(True)
        params = [nu, c]
        op = nu * ns.Laplacian() + c * ns.Convection()
    elif k == "burgers_learned_scalar":               # 0-dim tensors: one learnable value shared by the batch
        nu = torch.tensor(0.02, dtype=dtype, device=device).requires_grad_(True)
        c = torch.tensor(-0.8, dtype=dtype, device=device).requires_grad_(True)
        params = [nu, c]
        op = nu * ns.Laplacian() + c * ns.Convection()
    elif k == "heat_learned_source":
        x, y = mesh.bc_mesh_grid()
        src = (torch.sin(TWO_PI * x) * torch.cos(2 * TWO_PI * y)).to(dtype).contiguous().requires_grad_(True)
        params = [src]
        op = 0.05 * ns.Laplacian() + ns.ExplicitSource(src)
    elif k == "ks":
        op = -ns.Laplacian() - ns.Biharmonic() - ns.KSConvection()
    elif k == "kdv":
        op = -ns.SpatialDerivative(0, 3) - 6 * ns.Convection()
    elif k == "ns2d":
        op = 0.01 * ns.Laplacian() - ns.VorticityConvection()
    elif k == "ns2d_forced":
        y = mesh.bc_mesh_grid()[1]
        force = -0.1 * ns.ImplicitSource() - ns.ExplicitSource(4.0 * torch.cos(4.0 * y))
        op = 0.01 * ns.Laplacian() - ns.VorticityConvection() + force
    elif k == "beta_plane":
        op = 0.01 * ns.Laplacian() - ns.VorticityConvection() + 0.5 * ns.SpatialDerivative(0, 1)
    elif k == "ns3d":
        op = 0.01 * ns.Laplacian() + ns.NSPressureConvection()
    elif k == "allen_cahn":
        op = 0.05 * ns.Laplacian() + ns.ImplicitSource(lambda u: u - u ** 3)
    elif k == "conservative":
        op = 0.01 * ns.Laplacian() - ns.ConservativeConvection()
    else:
        raise KeyError(k)
    name = case["integrator"]
    enum = ns.ETDRKIntegrator if name.startswith("ETDRK") else ns.SETDRKIntegrator if name.startswith("SETDRK") \
        else ns.RKIntegrator
    op.set_integrator(getattr(enum, name))
    return op, params


def inputs(case, dtype):
    """Smooth seeded initial field and cotangent, identical for the fixture generator and the tests."""
    g = torch.Generator().manual_seed(sum(map(ord, case["name"])))
    shape = [m[2] for m in case["mesh"]]
    u = torch.randn(BATCH, case["C"], *shape, dtype=torch.float64, generator=g)
    u_hat = torch.fft.fftn(u, dim=list(range(2, u.dim())))
    for a, n in enumerate(shape):                      # keep |k| <= 3 per axis
        f = torch.fft.fftfreq(n, 1.0 / n).abs()
        u_hat = u_hat * (f <= 3).to(u_hat.dtype).reshape([1, 1] + [n if i == a else 1 for i in range(len(shape))])
    u = torch.fft.ifftn(u_hat, dim=list(range(2, u.dim()))).real
    u = u / u.abs().amax()
    if case.get("rough"):
        u = u + case["rough"] * torch.randn(BATCH, case["C"], *shape, dtype=torch.float64, generator=g)
    w = torch.randn(BATCH, case["C"], *shape, dtype=torch.float64, generator=g)
    return u.to(dtype).contiguous(), w.to(dtype).contiguous()


def run(ns, case, dtype, device="cpu"):
    """-> dict(y, grad_y, z, grad_z): y = integrate(u0, STEPS), z = op(u0); gradients of <y, w> and <z, w> w.r.t. u0."""
    mesh_info = [tuple(m) for m in case["mesh"]]
    u0, w = inputs(case, dtype)
    u0, w = u0.to(device), w.to(device)
    mesh = ns.MeshGrid(mesh_info, dtype=dtype, device=device)
    op, params = build(ns, case, mesh, dtype, device)
    x = u0.clone().requires_grad_(True)
    y = op.integrate(x, mesh=mesh, dt=case["dt"], step=STEPS)
    (y * w).sum().backward()
    out = {"y": y.detach(), "grad_y": x.grad}
    for i, p in enumerate(params):
        out[f"grad_y_p{i}"] = p.grad.detach().clone()
    if params:      # the reference caches L (and its graph) in the operator: a fresh operator per differentiated call
        op, params = build(ns, case, mesh, dtype, device)
    x2 = u0.clone().requires_grad_(True)
    z = op(x2, mesh=mesh)
    (z * w).sum().backward()
    out.update({"z": z.detach(), "grad_z": x2.grad})
    for i, p in enumerate(params):
        out[f"grad_z_p{i}"] = p.grad.detach().clone()
    if params:      # parameters alone: the state does not require grad
        op, params = build(ns, case, mesh, dtype, device)
        (op.integrate(u0.clone(), mesh=mesh, dt=case["dt"], step=STEPS) * w).sum().backward()
        for i, p in enumerate(params):
            out[f"grad_only_p{i}"] = p.grad.detach().clone()
    return out


def load(name, dtype):
    tag = "f32" if dtype == torch.float32 else "f64"
    with np.load(os.path.join(GRAD_DIR, f"{name}_{tag}.npz")) as z:
        out = {k: torch.from_numpy(z[k]) for k in z.files if k != "spec"}
        out["spec"] = json.loads(str(z["spec"]))
    return out
