"""The checks of the ops fixtures (tests/golden_ops, generated from the unmodified reference by
tests/golden/make_golden_ops.py), shared by the emulator run on CPU and the CUDA run on the GPU."""
import torch

from golden_util import rel_l2
from ops_util import build_operator, case_sources, load_ops

TOL = {"f32": 1e-5, "f64": 1e-12}


def _dtype(g):
    return torch.float32 if g["spec"]["dtype"] == "float32" else torch.float64


def check_ops_case(name, device):
    import torchfsm_b200 as fsm
    g = load_ops(name)
    spec, dtype = g["spec"], _dtype(g)
    tol = TOL[name[-3:]]
    mesh = fsm.MeshGrid([tuple(m) for m in spec["mesh"]], device=device, dtype=dtype)
    u0 = torch.from_numpy(g["u0"]).to(device)
    srcs = case_sources(spec, dtype, device)
    if spec["mode"] == "call":
        op = build_operator(fsm, spec["terms"], srcs, dtype, device)
        y = op(u0, mesh=mesh)
        assert y.shape == g["y"].shape
        assert rel_l2(y.cpu().numpy(), g["y"]) <= 10 * tol, name
        # the Fourier-space return is the Hermitian projection of the reference's (same field after .real(ifft))
        y_hat = op(u0, return_in_fourier=True)
        back = torch.fft.ifftn(y_hat, dim=tuple(range(2, y_hat.dim()))).real
        assert rel_l2(back.cpu().numpy(), g["y"]) <= 10 * tol, name
        # spectral input
        y2 = op(u_fft=torch.fft.fftn(u0, dim=tuple(range(2, u0.dim()))))
        assert rel_l2(y2.cpu().numpy(), g["y"]) <= 10 * tol, name
    elif spec["mode"] == "run_operators":
        ops = [build_operator(fsm, t, srcs, dtype, device) for t in spec["operators"]]
        outs = list(fsm.run_operators(u0, ops, mesh))
        assert len(outs) == len(ops)
        for i, y in enumerate(outs):
            assert rel_l2(y.cpu().numpy(), g[f"y{i}"]) <= 10 * tol, (name, i)
    elif spec["mode"] == "solve":
        op = build_operator(fsm, spec["terms"], srcs, dtype, device)
        y = op.solve(b=u0, mesh=mesh, n_channel=spec["C"])
        assert rel_l2(y.cpu().numpy(), g["y"]) <= 10 * tol, name
    else:
        from product_util import integrator_enum
        op = build_operator(fsm, spec["terms"], srcs, dtype, device)
        op.set_integrator(integrator_enum(spec["integrator"]))
        if name.endswith("f32") and torch.device(device).type != "cpu":
            # the plain-ETDRK fp32 tables cancel catastrophically, so tables built on the GPU differ from
            # the CPU-built ones of the fixture by more than the tolerance. Same inputs means the same tables: build
            # them with the same expressions where the fixture did (CPU) and hand them to the plan.
            from torchfsm_b200.integrator import build_tables, integrator_name
            m, c = op._pre_check(u0, None, mesh)
            op.register_mesh(m, c)
            L = op._state_dict["linear_coef"]
            if L is None:
                L = torch.zeros([1] * (u0.dim()), dtype=m.cdtype, device=device)
            lo = op._lowered
            iname = integrator_name(op._integrator, lo["program"] == 0 and lo["source_hat"] is None and not lo["external"])
            tabs = {k: v.to(device) for k, v in build_tables(iname, spec["dt"], L.cpu(), **op._integrator_config).items()}
            st = op._build_integrator(spec["dt"], u0.shape[0], tables=tabs)
            u_hat = st.step_half(st.r2c(u0), 1)
            u1 = st.c2r(u_hat)
            uT = st.c2r(st.step_half(u_hat, spec["steps"] - 1))
            assert rel_l2(u1.cpu().numpy(), g["u1"]) <= tol, name
            assert rel_l2(uT.cpu().numpy(), g["uT"]) <= tol, name
        u1 = op.integrate(u0, mesh=mesh, dt=spec["dt"], step=1)
        loose = 10 if (name.endswith("f32") and torch.device(device).type != "cpu") else 1    # self-built tables
        assert rel_l2(u1.cpu().numpy(), g["u1"]) <= tol * loose, name
        uT = op.integrate(u0, dt=spec["dt"], step=spec["steps"])
        assert rel_l2(uT.cpu().numpy(), g["uT"]) <= tol * loose * (spec["steps"] if name.endswith("f32") else 1), name
        assert rel_l2(op(u0).cpu().numpy(), g["rhs0"]) <= 10 * tol, name
