"""GPU parity tests (run with -m gpu on a B200). Everything goes through the public plugin layer
and the C ABI of libfsm_b200.so; the numpy oracle and the golden fixtures are the checkers.
Nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

from golden_util import golden_names, load_golden, golden_tables, rel_l2
from product_util import product_from_golden, product_operator

pytestmark = pytest.mark.gpu
TOL = {"f32": 1e-5, "f64": 1e-12}
SUPPORTED = golden_names()


@pytest.fixture(scope="module", autouse=True)
def cuda_library():
    from torchfsm_b200 import _cabi
    _cabi._lib = None
    lib = _cabi.lib()
    assert lib.fsm_backend() == 0, "GPU tests must run on the CUDA build, not the emulator"
    assert torch.cuda.is_available()
    yield


@pytest.mark.parametrize("name", SUPPORTED)
def test_cuda_matches_reference_golden(name):
    g = load_golden(name)
    spec, tol = g["spec"], TOL[name[-3:]]
    op, mesh, u0 = product_from_golden(g, "cuda")
    u1 = op.integrate(u0, mesh=mesh, dt=spec["dt"], step=1)
    assert rel_l2(u1.cpu().numpy(), g["u1"]) <= tol
    uT = op.integrate(u0, dt=spec["dt"], step=spec["steps"])
    # the bound is per step (north star); plain-ETDRK fp32 tables built on the GPU differ from the CPU-built
    # ones of the fixture through catastrophic cancellation, which adds up over the steps.
    # test_cuda_step_with_reference_tables below is the strict same-tables check.
    assert rel_l2(uT.cpu().numpy(), g["uT"]) <= tol * (spec["steps"] if name.endswith("f32") else 1)
    assert rel_l2(op(u0).cpu().numpy(), g["rhs0"]) <= 10 * tol
    st = op._state_dict["integrator"]
    full = st.half_to_full(st.r2c(u0))
    ref = torch.fft.fftn(torch.from_numpy(g["u0"]), dim=tuple(range(2, u0.dim())))
    assert rel_l2(full.cpu().numpy(), ref.numpy()) <= tol


@pytest.mark.parametrize("name", ["c3_ns2d_32_etdrk2_f32", "c2_ks2d_32_f32", "c4_burgers3d_16_f32"])
def test_cuda_step_with_reference_tables(name):
    """Same inputs INCLUDING the reference's own coefficient tables."""
    g = load_golden(name)
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cuda")
    m, c = op._pre_check(u0, None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(v).cuda() for k, v in golden_tables(g).items()}
    st = op._build_integrator(spec["dt"], u0.shape[0], tables=tabs)
    u_hat = st.step_half(st.r2c(u0), spec["steps"])
    assert rel_l2(st.c2r(u_hat).cpu().numpy(), g["uT"]) <= 1e-5


def _ns2d_terms(n, dtype):
    ax = torch.arange(n, dtype=dtype) * (2 * np.pi / n)
    y = ax.reshape(1, 1, 1, n).expand(1, 1, n, n)
    src = 4.0 * torch.cos(4.0 * y)
    return [("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {}), ("implicit_unit_source", -0.1, {}),
            ("explicit_source", -1, {"source": src})]


def _smooth(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(*shape, generator=g, dtype=dtype)
    dims = tuple(range(2, len(shape)))
    u_hat = torch.fft.fftn(u, dim=dims)
    for d in dims:
        n = shape[d]
        f = torch.fft.fftfreq(n, 1.0 / n).abs()
        keep = (f <= n // 8).to(u_hat.dtype)
        view = [1] * len(shape)
        view[d] = n
        u_hat = u_hat * keep.reshape(view)
    u = torch.fft.ifftn(u_hat, dim=dims).real
    return u / u.abs().amax(dim=tuple(range(1, len(shape))), keepdim=True)


def _conv(terms, fn):
    return [(k, c, {kk: (fn(vv) if isinstance(vv, torch.Tensor) else vv) for kk, vv in p.items()}) for k, c, p in terms]


@pytest.mark.parametrize("n,batch,dtype,tol", [(256, 3, torch.float32, 1e-5), (128, 2, torch.float64, 1e-12),
                                               (1024, 1, torch.float32, 1e-5)])
def test_ns2d_vs_oracle_seeded(n, batch, dtype, tol):
    """C3-shaped problem at sizes the oracle finishes in seconds; per-step parity over 5 steps."""
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    mesh_info = [(0, 2 * np.pi, n), (0, 2 * np.pi, n)]
    terms = _ns2d_terms(n, dtype)
    u0 = _smooth((batch, 1, n, n), dtype, seed=7)
    dt = 0.01
    ora = OracleOperator(_conv(terms, lambda t: t.numpy()))
    ora.register_mesh(mesh_info, 1, dtype="float32" if dtype == torch.float32 else "float64", workers=8)
    ora.set_integrator("ETDRK2")
    integ = ora.build_integrator(dt)
    op = product_operator(_conv(terms, lambda t: t.cuda()))
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    mesh = fsm.MeshGrid(mesh_info, device="cuda", dtype=dtype)
    m, c = op._pre_check(u0.cuda(), None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in integ.tables.items()}
    st = op._build_integrator(dt, batch, tables=tabs)
    u_hat = st.r2c(u0.cuda())
    ref_hat = ora.mesh.fft(u0.numpy())
    for step in range(5):
        u_hat = st.step_half(u_hat, 1)
        ref_hat = integ.step(ref_hat)
        got = st.c2r(u_hat).cpu().numpy()
        want = ora.mesh.ifft(ref_hat).real
        assert rel_l2(got, want) <= tol, f"step {step}"


@pytest.mark.parametrize("n,batch", [(128, 2)])
def test_burgers3d_vs_oracle_seeded(n, batch):
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    mesh_info = [(0, 1, n)] * 3
    terms = [("laplacian", 0.01, {}), ("convection", -1, {})]
    u0 = _smooth((batch, 3, n, n, n), torch.float32, seed=11)
    dt = 0.002
    ora = OracleOperator(terms).register_mesh(mesh_info, 3, dtype="float32", workers=8)
    ora.set_integrator("ETDRK2")
    integ = ora.build_integrator(dt)
    op = product_operator(terms)
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    mesh = fsm.MeshGrid(mesh_info, device="cuda", dtype=torch.float32)
    m, c = op._pre_check(u0.cuda(), None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in integ.tables.items()}
    st = op._build_integrator(dt, batch, tables=tabs)
    u_hat = st.step_half(st.r2c(u0.cuda()), 2)
    ref_hat = integ.step(integ.step(ora.mesh.fft(u0.numpy())))
    assert rel_l2(st.c2r(u_hat).cpu().numpy(), ora.mesh.ifft(ref_hat).real) <= 1e-5


# ---------------------------------------------------------------- full-size, size-independent properties
def test_c3_full_size_properties():
    """1024^2 x 64 (BASELINE config C3): transform round trip, Parseval, batch independence, Taylor-Green."""
    import torchfsm_b200 as fsm
    n, B = 1024, 64
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 2, device="cuda", dtype=torch.float32)
    x, y = mesh.bc_mesh_grid()
    op = fsm.pde.NavierStokesVorticity(Re=100, force=fsm.field.kolm_force(y))
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    u0 = fsm.field.diffused_noise(mesh, batch_size=B, generator=torch.Generator().manual_seed(0))
    assert u0.shape == (B, 1, n, n) and torch.isfinite(u0).all()
    uT = op.integrate(u0, mesh=mesh, dt=0.01, step=3)
    assert torch.isfinite(uT).all()
    st = op._state_dict["integrator"]
    back = st.c2r(st.r2c(u0))
    assert float((back - u0).norm() / u0.norm()) < 5e-7
    full = st.half_to_full(st.r2c(u0))
    assert abs(float((full.abs().double() ** 2).sum() / (n * n) / (u0.double() ** 2).sum()) - 1) < 1e-5
    del full, back
    uT2 = op.integrate(u0[:2].contiguous(), dt=0.01, step=3)
    assert float((uT2 - uT[:2]).norm() / uT[:2].norm()) < 1e-6
    tg = fsm.pde.NavierStokesVorticity(Re=50)
    tg.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    w0 = (2 * torch.cos(x) * torch.cos(y)).repeat(2, 1, 1, 1)
    wT = tg.integrate(w0, mesh=mesh, dt=0.05, step=10)
    assert float((wT - w0 * np.exp(-2 * 0.5 / 50)).abs().max()) < 2e-5


def test_c2_ks_batch_mean_property():
    """KS 256^2: the batch mean couples samples only through the k=0 bin."""
    import torchfsm_b200 as fsm
    n = 256
    mesh = fsm.MeshGrid([(0, 60, n)] * 2, device="cuda", dtype=torch.float32)
    u0 = torch.randn(8, 1, n, n, generator=torch.Generator().manual_seed(0)).cuda()
    both = fsm.pde.KuramotoSivashinskyHighDim().integrate(u0, mesh=mesh, dt=0.1, step=3)
    alone = fsm.pde.KuramotoSivashinskyHighDim().integrate(u0[:1].contiguous(), mesh=mesh, dt=0.1, step=3)
    diff = (both[:1] - alone).double()
    assert float(diff.std()) < 1e-5 * float(alone.abs().max()) and torch.isfinite(both).all()


def test_pure_diffusion_is_exact_on_gpu():
    import torchfsm_b200 as fsm
    n, nu, dt, steps = 512, 0.03, 0.1, 7
    mesh = fsm.MeshGrid([(0, 1, n), (0, 1, n)], device="cuda", dtype=torch.float64)
    x, y = mesh.bc_mesh_grid()
    u0 = torch.sin(2 * np.pi * 3 * x) * torch.cos(2 * np.pi * 2 * y)
    uT = (nu * fsm.Laplacian()).integrate(u0, mesh=mesh, dt=dt, step=steps)
    want = np.exp(-nu * (2 * np.pi) ** 2 * 13 * dt * steps) * u0
    assert float((uT - want).abs().max()) < 1e-12


_ROLE_SHAPES = [(8, 256, 8), (256, 8, 8), (8, 8, 256), (8, 512, 8), (512, 8, 8), (8, 8, 512),
                (512, 8), (8, 512), (256, 16), (16, 256)]


def test_fp64_shared_memory_limits_are_refused_up_front():
    """fp64 line buffers of 1024-point lines (and of 512-point last-axis lines of 3-D convection) exceed an SM's
    shared memory: the host layer says so instead of failing at launch."""
    import torchfsm_b200 as fsm
    for shape in [(1024, 16), (16, 1024), (8, 8, 512)]:
        nd = len(shape)
        mesh = fsm.MeshGrid([(0.0, 1.0, n) for n in shape], device="cuda", dtype=torch.float64)
        u0 = torch.zeros((1, nd) + shape, dtype=torch.float64, device="cuda")
        with pytest.raises(NotImplementedError):
            fsm.pde.Burgers(0.01).integrate(u0, mesh=mesh, dt=1e-3, step=1)


@pytest.mark.parametrize("shape,dtype,tol", [(s, torch.float32, 1e-5) for s in _ROLE_SHAPES] +
                         [(s, torch.float64, 1e-12) for s in _ROLE_SHAPES if s != (8, 8, 512)])
def test_long_lines_in_every_axis_role_vs_oracle(shape, dtype, tol):
    """The 256/512-point decompositions of C4/C5 (and their 2-D counterparts) in each role (x, middle, last
    axis) on thin grids the oracle finishes instantly: transforms vs torch.fft, one Burgers ETDRK2 step vs the oracle."""
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    nd = len(shape)
    mesh_info = [(0.0, 1.0 + 0.5 * i, n) for i, n in enumerate(shape)]
    u0 = _smooth((2, nd) + tuple(shape), torch.float64, seed=sum(shape)).to(dtype)
    terms = [("laplacian", 0.01, {}), ("convection", -1, {})]
    dt = 1e-4
    ora = OracleOperator(terms).register_mesh(mesh_info, nd, dtype="float32" if dtype == torch.float32 else "float64")
    ora.set_integrator("ETDRK2")
    integ = ora.build_integrator(dt)
    op = product_operator(terms)
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    mesh = fsm.MeshGrid(mesh_info, device="cuda", dtype=dtype)
    m, c = op._pre_check(u0.cuda(), None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in integ.tables.items()}
    st = op._build_integrator(dt, u0.shape[0], tables=tabs)
    u_hat = st.r2c(u0.cuda())
    dims = tuple(range(2, 2 + nd))
    assert rel_l2(st.half_to_full(u_hat).cpu().numpy(), torch.fft.fftn(u0.double(), dim=dims).numpy()) <= tol
    assert rel_l2(st.c2r(u_hat).cpu().numpy(), u0.numpy()) <= tol
    got = st.c2r(st.step_half(u_hat, 1)).cpu().numpy()
    want = ora.mesh.ifft(integ.step(ora.mesh.fft(u0.numpy()))).real
    assert rel_l2(got, want) <= tol


def test_full_run_drift_stays_small_on_gpu():
    """North star: drift over the full run. C1 at full size and C3 bounded to 256^2 x 2, 200 steps each, product vs
    oracle from the same state and tables. Chaotic growth is slow on these runs; the bound is generous on purpose."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import trajectory_drift as td
    dev = torch.device("cuda", 0)
    marks = [1, 50, 200]
    x = (np.arange(128) / 128.0).astype(np.float32).reshape(1, 1, 128)
    r1 = td.drift("C1", [(0.0, 1.0, 128)], lambda cv: [("laplacian", 0.01, {}), ("convection", -1, {})], "SETDRK4",
                  0.01, 200, (np.sin(2 * np.pi * x) + 0.5).astype(np.float32), torch.float32, dev, marks)
    assert max(r1["rel_l2_after_steps"].values()) <= 1e-4, r1
    n = 256
    ax = (np.arange(n) * (2 * np.pi / n)).astype(np.float32)
    src = (4.0 * np.cos(4.0 * ax)).reshape(1, 1, 1, n).repeat(n, axis=2).astype(np.float32)
    u0 = _smooth((2, 1, n, n), torch.float32, seed=3).numpy()
    r3 = td.drift("C3 bounded", [(0, 2 * np.pi, n)] * 2,
                  lambda cv: [("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {}),
                              ("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": cv(src)})],
                  "ETDRK2", 0.01, 200, u0, torch.float32, dev, marks)
    assert max(r3["rel_l2_after_steps"].values()) <= 1e-4, r3


# ---------------------------------------------------------------- round 2: the configs at size, per step, vs the oracle
def _per_step_vs_oracle(mesh_info, terms, n_channel, integrator, dt, u0, steps, tol, device="cuda", workers=8):
    """Product (CUDA library, oracle-built tables injected) against the oracle after EVERY step."""
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    from product_util import integrator_enum
    dtype = u0.dtype
    ora = OracleOperator(_conv(terms, lambda t: t.numpy()))
    ora.register_mesh(mesh_info, n_channel, dtype="float32" if dtype == torch.float32 else "float64", workers=workers)
    ora.set_integrator(integrator)
    integ = ora.build_integrator(dt)
    op = product_operator(_conv(terms, lambda t: t.to(device)))
    op.set_integrator(integrator_enum(integrator))
    mesh = fsm.MeshGrid(mesh_info, device=device, dtype=dtype)
    m, c = op._pre_check(u0.to(device), None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in integ.tables.items()}
    st = op._build_integrator(dt, u0.shape[0], tables=tabs)
    u_hat = st.r2c(u0.to(device))
    ref_hat = ora.mesh.fft(u0.numpy())
    errs = []
    for step in range(steps):
        u_hat = st.step_half(u_hat, 1)
        ref_hat = integ.step(ref_hat)
        errs.append(rel_l2(st.c2r(u_hat).cpu().numpy(), ora.mesh.ifft(ref_hat).real))
        assert errs[-1] <= tol, f"step {step}: {errs}"
    return errs


def _taylor_green(n, dtype, amp=0.05):
    ax = torch.arange(n, dtype=dtype) * (2 * np.pi / n)
    x, y, z = ax.reshape(1, 1, n, 1, 1), ax.reshape(1, 1, 1, n, 1), ax.reshape(1, 1, 1, 1, n)
    u = torch.cat([torch.sin(x) * torch.cos(y) * torch.cos(z), -torch.cos(x) * torch.sin(y) * torch.cos(z),
                   torch.zeros(1, 1, n, n, n, dtype=dtype)], dim=1)
    return u + amp * _smooth((1, 3, n, n, n), dtype, seed=5)


@pytest.mark.parametrize("n", [64, 128])
def test_c5_ns3d_setdrk4_vs_oracle(n):
    """C5's program and integrator (NS velocity form with pressure projection, SETDRK4, 2/3 dealiasing) at 64^3 and
    128^3: Taylor-Green + noise, three steps, each within 1e-5 of the oracle (_navier_stokes.py:231-254,
    _setdrk_step.py:55-82)."""
    terms = [("ns_pressure_convection", 1, {}), ("laplacian", 1 / 1600, {})]
    _per_step_vs_oracle([(0, 2 * np.pi, n)] * 3, terms, 3, "SETDRK4", 0.01, _taylor_green(n, torch.float32), 3, 1e-5)


def test_c5_ns3d_setdrk4_fp64_vs_oracle():
    terms = [("ns_pressure_convection", 1, {}), ("laplacian", 1 / 1600, {})]
    _per_step_vs_oracle([(0, 2 * np.pi, 64)] * 3, terms, 3, "SETDRK4", 0.01, _taylor_green(64, torch.float64), 3, 1e-12)


def test_c4_burgers3d_setdrk4_vs_oracle():
    """C4's program and integrator at 128^3 x 2 (the fixture-sized cases stop at 16^3)."""
    terms = [("laplacian", 0.01, {}), ("convection", -1, {})]
    u0 = _smooth((2, 3, 128, 128, 128), torch.float32, seed=11)
    _per_step_vs_oracle([(0, 1, 128)] * 3, terms, 3, "SETDRK4", 0.002, u0, 3, 1e-5)


def test_c2_ks2d_setdrk4_vs_oracle_at_size():
    """C2 at its own grid (256^2, L = 60) with SETDRK4 and a batch of 8: per-step parity, including the batch mean
    that couples the samples through the k = 0 bin (_ks_convection.py:34-36)."""
    terms = [("laplacian", -1, {}), ("biharmonic", -1, {}), ("ks_convection", -1, {})]
    u0 = torch.randn(8, 1, 256, 256, generator=torch.Generator().manual_seed(0))
    _per_step_vs_oracle([(0, 60, 256)] * 2, terms, 1, "SETDRK4", 0.5, u0, 3, 1e-5)


def test_plan_follows_the_tensor_device_not_the_current_device():
    """A mesh on cuda:1 while cuda:0 is current: tables, kernels and results must live on cuda:1."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = load_golden("c3_ns2d_32_etdrk2_f32")
    torch.cuda.set_device(0)
    op, mesh, u0 = product_from_golden(g, "cuda:1")
    uT = op.integrate(u0, mesh=mesh, dt=g["spec"]["dt"], step=g["spec"]["steps"])
    assert uT.device == torch.device("cuda", 1) and torch.cuda.current_device() == 0
    assert rel_l2(uT.cpu().numpy(), g["uT"]) <= 1e-5 * g["spec"]["steps"]


def test_integrate_stream_equals_integrate():
    """The pipelined host entry point (three streams, double-buffered staging) returns exactly what integrate() does."""
    import torchfsm_b200 as fsm
    n, B = 256, 4
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 2, device="cuda", dtype=torch.float32)
    _, y = mesh.bc_mesh_grid()
    op = fsm.pde.NavierStokesVorticity(Re=100, force=fsm.field.kolm_force(y))
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    batches = [_smooth((B, 1, n, n), torch.float32, seed=s).pin_memory() for s in range(5)]
    outs = op.integrate_stream(batches, dt=0.01, step=3, mesh=mesh)
    torch.cuda.synchronize()
    for b, o in zip(batches, outs):
        want = op.integrate(b.cuda(), dt=0.01, step=3)
        assert o.device.type == "cpu" and torch.equal(o, want.cpu())


def test_recorders_and_solve_on_gpu():
    """§8(f2)/(f4): physical frames straight from the C2R pass (AutoRecorder, CPURecorder) and LinearOperator.solve
    on the CUDA library (traj_recorder.py:95-148, operator/_base.py:217-262)."""
    import torchfsm_b200 as fsm
    g = load_golden("c3_ns2d_32_etdrk2_f32")
    op, mesh, u0 = product_from_golden(g, "cuda")
    dt, steps = g["spec"]["dt"], g["spec"]["steps"]
    final = op.integrate(u0, mesh=mesh, dt=dt, step=steps)
    for rec in (fsm.AutoRecorder(), fsm.CPURecorder()):
        traj = op.integrate(u0, dt=dt, step=steps, trajectory_recorder=rec)
        assert traj.shape == (u0.shape[0], steps + 1) + tuple(u0.shape[1:])
        assert float((traj[:, 0].to(u0.device) - u0).abs().max()) < 1e-5
        assert float((traj[:, -1].to(u0.device) - final).abs().max()) < 1e-6
    n = 64
    mesh2 = fsm.MeshGrid([(0, 2 * np.pi, n)] * 2, device="cuda", dtype=torch.float64)
    x, y = mesh2.bc_mesh_grid()
    phi = torch.sin(3 * x) * torch.cos(2 * y)
    rhs = -13.0 * phi                                           # lap(phi)
    sol = fsm.Laplacian().solve(b=rhs, mesh=mesh2, n_channel=1)
    assert float((sol - phi).abs().max()) < 1e-12


def test_genuine_reference_operator_on_the_gpu_dropin():
    """the unmodified reference (baseline/_ref, its cuFFT + ATen path) against the same reference
    operator object stepping through the fused kernels (torchfsm_b200.reference_adapter) on this GPU, same inputs."""
    from test_reference_adapter import _reference, run_dropin
    torchfsm = _reference()
    if torchfsm is None:
        pytest.skip("reference not importable on this box (baseline/_ref did not travel)")
    run_dropin(torchfsm, "cuda", torch.float32, 1e-5)
    run_dropin(torchfsm, "cuda", torch.float64, 1e-12)


# ---------------------------------------------------------------- round 2: the operators around the hot path on the GPU
def _ops_names():
    from ops_util import ops_names
    return ops_names()


@pytest.mark.parametrize("name", _ops_names())
def test_cuda_ops_match_reference(name):
    """Grad/Div/Curl, Vorticity2Velocity, the pressure diagnostics, ConservativeConvection, ImplicitSource(func),
    NSPressureConvection with an external force and in 2-D, per-sample coefficients, linear operators with a source,
    solve and run_operators on the CUDA library against vectors of the unmodified reference (tests/golden_ops)."""
    from ops_checks import check_ops_case
    check_ops_case(name, "cuda")


def test_batched_viscosity_at_size_vs_per_sample_runs():
    """Per-sample tables at C3's grid: a batch with three viscosities equals three single-viscosity runs."""
    import torchfsm_b200 as fsm
    n = 256
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 2, device="cuda", dtype=torch.float32)
    u0 = _smooth((3, 1, n, n), torch.float32, seed=2).cuda()
    nus = [0.01, 0.002, 0.05]
    nu = torch.tensor(nus, device="cuda").reshape(3, 1, 1, 1)
    op = nu * fsm.Laplacian() - fsm.VorticityConvection()
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    got = op.integrate(u0, mesh=mesh, dt=0.01, step=4)
    for i, v in enumerate(nus):
        one = v * fsm.Laplacian() - fsm.VorticityConvection()
        one.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
        want = one.integrate(u0[i:i + 1].contiguous(), mesh=mesh, dt=0.01, step=4)
        assert float((got[i:i + 1] - want).norm() / want.norm()) < 1e-6


# ---------------------------------------------------------------- slab decomposition at size on ONE GPU
def _slab_ranks_in_one_process(P, n, steps, integrator, nsub=1):
    """Every rank's plan of a P-way slab decomposition lives in this process on one GPU; the all-to-all between the
    phases is done by copying the rank blocks between the plans' exchange buffers (exactly what all_to_all_single /
    the copy engines do between GPUs). Returns (slab result assembled over the ranks, single-plan result)."""
    import torchfsm_b200 as fsm
    dev = torch.device("cuda", 0)
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 3, device=dev, dtype=torch.float32)
    u0 = _taylor_green(n, torch.float32).to(dev)
    dt = 0.01

    def make(rank=None):
        op = fsm.pde.NavierStokes(Re=1600)
        op.set_integrator(integrator)
        if rank is not None:
            op.set_slab_decomposition(group=None, rank=rank, nranks=P, nsub=nsub, exchange="nccl", graph=False)
        nxl = n // P
        u = u0 if rank is None else u0[:, :, rank * nxl:(rank + 1) * nxl].contiguous()
        m, c = op._pre_check(u, None, mesh)
        op.register_mesh(m, c)
        return op, op._build_integrator(dt, 1), u

    _, st1, _ = make()
    want = st1.c2r(st1.step_half(st1.r2c(u0), steps))
    ranks = [make(r) for r in range(P)]
    sts = [st for _, st, _ in ranks]

    def exchange(which, count, offset=0):
        blk = count // P
        for r in range(P):            # receive block q of rank r <- send block r of rank q
            for q in range(P):
                sts[r]._recv[which][offset + q * blk: offset + (q + 1) * blk].copy_(
                    sts[q]._send[which][offset + r * blk: offset + (r + 1) * blk])

    def phase(op, stage, ph, states, auxs, snd, rcv, sub=0, H=1):
        for st, x, a in zip(sts, states, auxs):
            st._slab_phase(op, stage, ph, x, a, snd, rcv, sub, H)

    # forward transform of the local slabs
    hats = [st.empty_half() for st in sts]
    phys = [u for _, _, u in ranks]
    c2 = sts[0]._slab_counts[2][1]
    phase(2, 0, 1, hats, phys, 1, None)
    exchange(1, c2)
    phase(2, 0, 2, hats, phys, None, 1)
    none = [None] * P
    H = sts[0].nsub
    c1s, c2s = sts[0]._slab_counts[0]
    for _ in range(steps):
        for stage in range(sts[0].n_stages):
            phase(0, stage, 0, hats, none, 0, None, 0, H)
            for h in range(H):
                exchange(0, c1s // H, h * (c1s // H))
                phase(0, stage, 1, hats, none, 1, 0, h, H)
                exchange(1, c2s // H, h * (c2s // H))
            phase(0, stage, 2, hats, none, None, 1, 0, H)
    outs = [torch.empty_like(u) for u in phys]
    c1 = sts[0]._slab_counts[3][0]
    phase(3, 0, 0, hats, outs, 0, None)
    exchange(0, c1)
    phase(3, 0, 1, hats, outs, None, 0)
    return torch.cat(outs, dim=2), want


@pytest.mark.parametrize("P,nsub", [(2, 1), (4, 2), (8, 1)])
def test_slab_decomposition_equals_single_plan_at_128cubed(P, nsub):
    """C5's path at 128^3 (NS velocity form, SETDRK4, 2/3 dealiasing): the P-way slab-decomposed plans (cyclic ky
    ownership, kept lines only on the inverse exchange, sub-slabs) against the single plan of the same grid, 2 steps."""
    import torchfsm_b200 as fsm
    got, want = _slab_ranks_in_one_process(P, 128, 2, fsm.SETDRKIntegrator.SETDRK4, nsub=nsub)
    assert float((got - want).norm() / want.norm()) <= 1e-6


def test_linear_operators_and_maps_are_differentiable_on_gpu():
    """Adjoint identity <A v, g> = <v, A^T g> for the differentiable (linear) part of the path on the CUDA library, and
    the gradient of a diffusion run against the closed form (the step is self-adjoint)."""
    import torchfsm_b200 as fsm
    torch.manual_seed(0)
    mesh3 = fsm.MeshGrid([(0, 1.0, 32), (0, 2.0, 16), (0, 1.5, 32)], device="cuda", dtype=torch.float64)
    mesh2 = fsm.MeshGrid([(0, 1.0, 64), (0, 2.0, 128)], device="cuda", dtype=torch.float64)
    for op, mesh, c in [(fsm.Curl(), mesh3, 3), (fsm.Div(), mesh2, 2), (fsm.Vorticity2Velocity(), mesh2, 1),
                        (0.3 * fsm.Laplacian() + 0.2 * fsm.SpatialDerivative(1, 3), mesh2, 1)]:
        shape = (2, c) + tuple(m[2] for m in mesh.mesh_info)
        u = torch.randn(*shape, dtype=torch.float64, device="cuda", requires_grad=True)
        y = op(u, mesh=mesh)
        g = torch.randn_like(y)
        (gu,) = torch.autograd.grad(y, u, g)
        v = torch.randn(*shape, dtype=torch.float64, device="cuda")
        lhs, rhs = float((op(v, mesh=mesh) * g).sum()), float((v * gu).sum())
        assert abs(lhs - rhs) <= 1e-10 * max(1.0, abs(lhs))
    u = torch.randn(2, 1, 64, 128, dtype=torch.float64, device="cuda", requires_grad=True)
    out = (0.05 * fsm.Laplacian()).integrate(u, mesh=mesh2, dt=0.1, step=3)
    (gu,) = torch.autograd.grad(out.sum(), u)
    assert float((gu - 1.0).abs().max()) < 1e-12          # d/du sum(exp(L t) u) = exp(L t)^T 1 = 1 (the mean mode is kept)


@pytest.mark.gpu
def test_functional_forms_and_batch_wise_recorders_on_gpu(tmp_path):
    """functional.py, DiskRecorder and RandomBatchWisedRecorder through the CUDA library (the CPU twin of this test,
    tests/test_functional_recorders.py, also checks them against the unmodified reference)."""
    import numpy as np
    import torchfsm_b200 as fsm
    import torchfsm_b200.functional as F
    dev = "cuda"
    mesh = fsm.MeshGrid([(0, 6.28, 64), (0, 3.0, 128)], device=dev, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    w = torch.randn(2, 1, 64, 128, generator=g).to(dev)
    w = (0.002 * fsm.Laplacian()).integrate(w, mesh=mesh, dt=1.0, step=1)
    assert float((F.grad(w, mesh=mesh) - fsm.Grad()(w, mesh=mesh)).abs().max()) == 0.0
    assert float((F.vorticity2velocity(w, mesh=mesh) - fsm.Vorticity2Velocity()(w, mesh=mesh)).abs().max()) == 0.0
    p = F.vorticity2pressure(w, mesh=mesh, external_force=-0.1 * fsm.ImplicitSource())
    assert p.shape == (2, 1, 64, 128) and torch.isfinite(p).all()
    op = 0.01 * fsm.Laplacian() - fsm.VorticityConvection()
    whole = op.integrate(w, mesh=mesh, dt=0.01, step=12, trajectory_recorder=fsm.AutoRecorder())
    rec = fsm.DiskRecorder(cache_dir=str(tmp_path) + "/", cache_freq=5)
    assert op.integrate(w, mesh=mesh, dt=0.01, step=12, trajectory_recorder=rec) is None
    parts = [torch.load(f) for f in rec.files]
    assert [q.shape[1] for q in parts] == [5, 5, 3]
    assert float((torch.cat(parts, dim=1) - whole.cpu()).abs().max()) == 0.0
    np.random.seed(5)
    rnd = fsm.RandomBatchWisedRecorder(simulation_steps=12, recorder_interval=2, n_recorded_frames=2)
    traj = op.integrate(w, mesh=mesh, dt=0.01, step=12, trajectory_recorder=rnd)
    ids = rnd._recorded_frame_id
    for b in range(2):
        for j in range(2):
            assert float((traj[b, j] - whole[b, ids[b, j]]).abs().max()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_fourier_mesh_transforms_on_gpu(dtype):
    """FourierMesh.fft / .ifft (the reference's transform choke point, mesh.py:481-491) on the CUDA library's passes
    against cuFFT, for real fields, complex fields and non-Hermitian spectra."""
    import torchfsm_b200 as fsm
    tol = 2e-5 if dtype == torch.float32 else 1e-12
    cd = torch.complex64 if dtype == torch.float32 else torch.complex128
    g = torch.Generator().manual_seed(9)
    for mesh_info in ([(0, 1, 256)], [(0, 1, 128), (0, 2, 256)], [(0, 1, 32), (0, 2, 64), (0, 3, 32)]):
        f = fsm.FourierMesh(mesh_info, device="cuda", dtype=dtype)
        shape = [m[2] for m in mesh_info]
        dims = list(range(-len(shape), 0))
        u = torch.randn(2, 3, *shape, dtype=dtype, generator=g).cuda()
        spec = torch.randn(2, 3, *shape, dtype=cd, generator=g).cuda()
        scale = float(torch.fft.fftn(u, dim=dims).abs().max())
        assert float((f.fft(u) - torch.fft.fftn(u, dim=dims)).abs().max()) < tol * scale
        assert float((f.fft(spec) - torch.fft.fftn(spec, dim=dims)).abs().max()) < tol * scale * 2
        assert float((f.ifft(spec) - torch.fft.ifftn(spec, dim=dims)).abs().max()) < tol
        assert float((f.ifft(f.fft(u)).real - u).abs().max()) < tol * 10


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_user_defined_cores_on_gpu(dtype):
    """The reference's extension protocol (LinearCoef / NonlinearFunc / CoreGenerator, operator/_base.py:16-131) through
    the CUDA library: the user core's ``f_mesh.fft`` / ``f_mesh.ifft`` are the library's passes. Checked against the same
    physics written with built-in operators (tests/test_custom_cores.py checks the reference itself on CPU)."""
    import torchfsm_b200 as fsm
    from test_custom_cores import make_cores
    Cubic, GradientSquared, HyperViscosity, ByChannels = make_cores(fsm)
    tol = 2e-5 if dtype == torch.float32 else 1e-12
    mesh = fsm.MeshGrid([(0, 1, 128), (0, 1, 256)], device="cuda", dtype=dtype)
    g = torch.Generator().manual_seed(17)
    u0 = torch.randn(2, 1, 128, 256, dtype=dtype, generator=g).cuda()
    u0 = (0.0005 * fsm.Laplacian()).integrate(u0, mesh=mesh, dt=1.0, step=1)
    u0 = u0 / u0.abs().max()
    builtin = 0.05 * fsm.Laplacian() - 1e-6 * fsm.Biharmonic() + fsm.ImplicitSource(lambda u: -u ** 3)
    custom = 0.05 * fsm.Laplacian() + fsm.LinearOperator(HyperViscosity(1e-6)) + fsm.Operator(ByChannels())
    for op in (builtin, custom):
        op.set_integrator(fsm.SETDRKIntegrator.SETDRK4)
    want = builtin.integrate(u0, mesh=mesh, dt=0.002, step=4)
    got = custom.integrate(u0, mesh=mesh, dt=0.002, step=4)
    assert float((got - want).norm() / want.norm()) < tol
    gs = fsm.NonlinearOperator(GradientSquared())(u0, mesh=mesh)
    gx, gy = fsm.Grad()(u0, mesh=mesh).unbind(dim=1)
    ref = gx * gx + gy * gy
    assert float((gs[:, 0] - ref).norm() / ref.norm()) < 10 * tol
