"""Pin the CPU oracle against golden vectors produced by the reference itself
(tests/golden/make_golden.py) and against analytic known answers."""
import numpy as np
import pytest

from golden_util import golden_names, load_golden, oracle_from_golden, golden_tables, rel_l2

TOL = {"f32": 1e-5, "f64": 1e-12}


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    spec = g["spec"]
    tol = TOL[name[-3:]]
    op = oracle_from_golden(g)
    if "linear_coef" in g:
        L = np.broadcast_to(op.linear_coef, g["linear_coef"].shape)
        assert rel_l2(L, g["linear_coef"]) <= tol
    u0_hat = op.mesh.fft(g["u0"])
    if "n0_hat" in g:
        assert rel_l2(op.nonlinear_func(u0_hat), g["n0_hat"]) <= 10 * tol  # compared in spectral space
    integ = op.build_integrator(spec["dt"])
    for k, ref in golden_tables(g).items():
        assert rel_l2(np.broadcast_to(integ.tables[k], ref.shape), ref) <= tol, k
    assert rel_l2(integ.step(u0_hat), g["u1_hat"]) <= tol
    u = u0_hat
    for _ in range(spec["steps"]):
        u = integ.step(u)
    assert rel_l2(op.mesh.ifft(u).real, g["uT"]) <= tol
    assert rel_l2(u, g["uT_hat"]) <= tol
    assert rel_l2(op(g["u0"]), g["rhs0"]) <= 10 * tol


@pytest.mark.parametrize("name", ["c3_ns2d_32_etdrk2_f32", "c2_ks2d_32_f32", "c4_burgers3d_16_f64"])
def test_oracle_integrate_driver(name):
    g = load_golden(name)
    spec = g["spec"]
    op = oracle_from_golden(g)
    uT = op.integrate(g["u0"], dt=spec["dt"], step=spec["steps"])
    assert rel_l2(uT, g["uT"]) <= TOL[name[-3:]]
    traj = op.integrate(g["u0"], dt=spec["dt"], step=2, record_every=1)
    assert traj.shape[1] == 3 and rel_l2(traj[:, 1], g["u1"]) <= TOL[name[-3:]]


def test_known_answers_derivatives():
    """docs/tutorials/derivative_example.ipynb cells 4-32: closed forms on a 3-D box."""
    from oracle import OracleMesh, OracleOperator
    mesh_info = [(0, 2 * np.pi, 32), (0, 4 * np.pi, 64), (0, 8 * np.pi, 128)]
    m = OracleMesh(mesh_info, "float64")
    ax = [np.arange(n) * (b - a) / n for (a, b, n) in mesh_info]
    x, y, z = np.meshgrid(*ax, indexing="ij")
    u = np.stack([np.sin(x), np.cos(y), np.sin(z) + np.cos(z)])[None]
    lap = OracleOperator([("laplacian", 1, {})]).register_mesh(m, 3)
    assert np.abs(lap(u) - np.stack([-np.sin(x), -np.cos(y), -np.sin(z) - np.cos(z)])[None]).max() < 1e-10
    conv = OracleOperator([("convection", 1, {})], de_aliasing_rate=1.0).register_mesh(m, 3)
    want = np.stack([np.sin(x) * np.cos(x), -np.sin(y) * np.cos(y), np.cos(z) ** 2 - np.sin(z) ** 2])[None]
    assert np.abs(conv(u) - want).max() < 1e-10


def test_known_answer_pure_diffusion_is_exact():
    """ETDRK0 is exact for linear terms: u_hat(t) = exp(-nu k^2 t) u_hat(0)."""
    from oracle import OracleOperator
    n, nu, dt, steps = 64, 0.03, 0.1, 7
    x = np.arange(n) / n
    u0 = (np.sin(2 * np.pi * 3 * x) + 0.25 * np.cos(2 * np.pi * 5 * x))[None, None]
    op = OracleOperator([("laplacian", nu, {})]).register_mesh([(0, 1, n)], 1, dtype="float64")
    uT = op.integrate(u0, dt=dt, step=steps)
    t = dt * steps
    want = (np.exp(-nu * (2 * np.pi * 3) ** 2 * t) * np.sin(2 * np.pi * 3 * x)
            + 0.25 * np.exp(-nu * (2 * np.pi * 5) ** 2 * t) * np.cos(2 * np.pi * 5 * x))
    assert np.abs(uT[0, 0] - want).max() < 1e-12


def test_known_answer_taylor_green_2d():
    """2-D Taylor-Green vortex: the convection term vanishes, omega decays as exp(-2 nu t)."""
    from oracle import OracleOperator
    n, Re, dt, steps = 32, 50.0, 0.05, 10
    ax = np.arange(n) * 2 * np.pi / n
    x, y = np.meshgrid(ax, ax, indexing="ij")
    w0 = (2 * np.cos(x) * np.cos(y))[None, None]
    op = OracleOperator([("vorticity_convection", -1, {}), ("laplacian", 1 / Re, {})])
    op.register_mesh([(0, 2 * np.pi, n)] * 2, 1, dtype="float64")
    op.set_integrator("ETDRK2")
    wT = op.integrate(w0, dt=dt, step=steps)
    assert np.abs(wT - w0 * np.exp(-2 * dt * steps / Re)).max() < 1e-10


def test_ks_batch_mean_only_shifts_dc():
    """the KS mean couples the batch only through the k=0 bin."""
    from oracle import OracleOperator
    g = load_golden("c2_ks2d_32_f64")
    spec = g["spec"]
    both = oracle_from_golden(g).integrate(g["u0"][:2], dt=spec["dt"], step=3)
    alone = oracle_from_golden(g).integrate(g["u0"][:1], dt=spec["dt"], step=3)
    diff = both[:1] - alone
    assert diff.std() < 1e-12 and abs(diff.mean()) > 1e-8


# ---------------------------------------------------------------- the operators around the path (tests/golden_ops)
def _oracle_terms(terms, spec, dtype):
    """ops_util term lists -> oracle term lists (numpy sources, numpy callables, nested force operators)."""
    import torch
    from ops_util import FUNCS, source_field
    from oracle import OracleOperator
    tdt = torch.float32 if dtype == "float32" else torch.float64
    out = []
    for kind, coef, params in terms:
        p = {k: v for k, v in dict(params).items() if k != "_ndim"}
        if "source" in p:
            p["source"] = source_field(p["source"], [tuple(m) for m in spec["mesh"]], tdt).numpy()
        if "func" in p:
            f = FUNCS[p["func"]]
            p["func"] = lambda u, f=f: f(torch.from_numpy(np.ascontiguousarray(u))).numpy()
        if p.get("force") is not None:
            p["force"] = OracleOperator(_oracle_terms(p["force"], spec, dtype))
        if isinstance(coef, (list, tuple)):
            coef = np.asarray(coef, dtype=dtype).reshape([len(coef), 1] + [1] * len(spec["mesh"]))
        out.append((kind, coef, p))
    return out


def _ops_names():
    from ops_util import ops_names
    return ops_names()


@pytest.mark.parametrize("name", _ops_names())
def test_oracle_matches_reference_ops(name):
    """The oracle's restatement of the cores around the path against vectors generated from the unmodified reference."""
    from ops_util import load_ops
    from oracle import OracleOperator
    g = load_ops(name)
    spec = g["spec"]
    tol = 1e-5 if name.endswith("f32") else 1e-12
    mesh_info = [tuple(m) for m in spec["mesh"]]
    u0 = g["u0"]

    def build(terms):
        op = OracleOperator(_oracle_terms(terms, spec, spec["dtype"]))
        op.register_mesh(mesh_info, spec["C"], dtype=spec["dtype"])
        return op
    if spec["mode"] == "call":
        assert rel_l2(build(spec["terms"])(u0.copy()), g["y"]) <= tol
    elif spec["mode"] == "run_operators":
        for i, terms in enumerate(spec["operators"]):
            assert rel_l2(build(terms)(u0.copy()), g[f"y{i}"]) <= tol
    elif spec["mode"] == "solve":
        op = build(spec["terms"])
        L = op.linear_coef
        inv = np.where(L == 0, 1.0, 1 / np.where(L == 0, 1, L))          # operator/_base.py:250-255
        assert rel_l2(op.mesh.ifft(op.mesh.fft(u0) * inv).real, g["y"]) <= tol
    else:
        op = build(spec["terms"])
        op.set_integrator(spec["integrator"])
        assert rel_l2(op.integrate(u0.copy(), dt=spec["dt"], step=1), g["u1"]) <= tol
        assert rel_l2(op.integrate(u0.copy(), dt=spec["dt"], step=spec["steps"]), g["uT"]) <= tol * (3 if name.endswith("f32") else 1)
        assert rel_l2(op(u0.copy()), g["rhs0"]) <= 10 * tol      # one evaluation of stiff symbols (k^4): as in the product tests
