"""CPU checks of the product's host logic and kernel logic.

The kernels of torchfsm_b200/csrc are compiled for the host by a fiber-based emulator
(tests/emu) and driven through the SAME C ABI and Python plugin layer as on the GPU, then
compared with the reference's golden vectors. This is test infrastructure: the product never
loads the emulator build (see test_product_boundary.py)."""
import numpy as np
import pytest
import torch

from golden_util import golden_names, load_golden, rel_l2
from product_util import build_emulator, product_from_golden, product_operator

TOL = {"f32": 1e-5, "f64": 1e-12}
SUPPORTED = golden_names()


@pytest.fixture(scope="module", autouse=True)
def emulator():
    from torchfsm_b200 import _cabi
    _cabi.use_library(build_emulator())
    assert _cabi.is_emulator()
    yield
    _cabi._lib = None


@pytest.mark.parametrize("name", SUPPORTED)
def test_emulated_kernels_match_reference_golden(name):
    g = load_golden(name)
    spec, tol = g["spec"], TOL[name[-3:]]
    op, mesh, u0 = product_from_golden(g, "cpu")
    u1 = op.integrate(u0, mesh=mesh, dt=spec["dt"], step=1)
    assert rel_l2(u1.numpy(), g["u1"]) <= tol
    uT = op.integrate(u0, dt=spec["dt"], step=spec["steps"])
    assert rel_l2(uT.numpy(), g["uT"]) <= tol
    uT_hat = op.integrate(u0, dt=spec["dt"], step=spec["steps"], return_in_fourier=True)
    assert uT_hat.shape == g["uT_hat"].shape
    # the drop-in returns the Hermitian projection of the reference's spectrum:
    # identical after .real(ifft(.))
    back = torch.fft.ifftn(uT_hat, dim=tuple(range(2, uT_hat.dim()))).real
    assert rel_l2(back.numpy(), g["uT"]) <= tol
    assert rel_l2(op(u0).numpy(), g["rhs0"]) <= 10 * tol


def test_transforms_roundtrip_and_full_spectrum():
    g = load_golden("burgers2d_32x64_etdrk2_f64")
    op, mesh, u0 = product_from_golden(g, "cpu")
    op.integrate(u0, mesh=mesh, dt=g["spec"]["dt"], step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u0)
    assert rel_l2(st.c2r(u_hat).numpy(), u0.numpy()) < 1e-14
    full = st.half_to_full(u_hat)
    ref = torch.fft.fftn(u0, dim=(2, 3))
    assert rel_l2(full.numpy(), ref.numpy()) < 1e-14
    assert rel_l2(st.full_to_half(ref).numpy(), u_hat.numpy()) < 1e-14


def test_u0_fft_input_and_recorder_protocol():
    import torchfsm_b200 as fsm
    g = load_golden("c3_ns2d_32_etdrk2_f64")
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cpu")
    u0_fft = torch.fft.fftn(u0, dim=(2, 3))
    uT = op.integrate(u_0_fft=u0_fft, mesh=mesh, dt=spec["dt"], step=spec["steps"])
    assert rel_l2(uT.numpy(), g["uT"]) <= 1e-12
    rec = fsm.AutoRecorder(fsm.IntervalController(interval=2))
    traj = op.integrate(u0, dt=spec["dt"], step=4, trajectory_recorder=rec)
    assert traj.shape == (u0.shape[0], 3) + tuple(u0.shape[1:])          # steps 0, 2, 4
    assert rel_l2(traj[:, 0].numpy(), g["u0"]) <= 1e-12


def test_stepper_speaks_the_integrator_protocol():
    """_state_dict['integrator'] exposes .dt/.step/.forward on full spectra (operator/_base.py:462-491)."""
    g = load_golden("ns2d_32_setdrk4_f64")
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cpu")
    op.integrate(u0, mesh=mesh, dt=spec["dt"], step=1)
    st = op._state_dict["integrator"]
    assert st.dt == spec["dt"]
    u1_full = st.forward(torch.fft.fftn(u0, dim=(2, 3)), spec["dt"])
    assert rel_l2(torch.fft.ifftn(u1_full, dim=(2, 3)).real.numpy(), g["u1"]) <= 1e-12


def test_reference_tables_can_be_injected():
    """The ETD tables are inputs of the C ABI: feeding the reference's own tables reproduces its step."""
    from golden_util import golden_tables
    g = load_golden("c3_ns2d_32_etdrk2_f32")
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cpu")
    m, c = op._pre_check(u0, None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(v) for k, v in golden_tables(g).items()}
    st = op._build_integrator(spec["dt"], u0.shape[0], tables=tabs)
    u_hat = st.step_half(st.r2c(u0), spec["steps"])
    assert rel_l2(st.c2r(u_hat).numpy(), g["uT"]) <= 1e-5


def test_unsupported_requests_fail_loudly():
    import torchfsm_b200 as fsm
    u = torch.zeros(1, 1, 24, 24, dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        (fsm.Laplacian() - fsm.KSConvection()).integrate(u, mesh=[(0, 1, 24), (0, 1, 24)], dt=0.1, step=1)
    u = torch.zeros(1, 2, 32, 32, dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        (fsm.Laplacian() - fsm.Convection() - fsm.KSConvection()).integrate(
            u, mesh=[(0, 1, 32), (0, 1, 32)], dt=0.1, step=1)
    with pytest.raises(NotImplementedError):     # a host-composed core cannot ride inside a fused convective program
        (fsm.Laplacian() - fsm.Convection() + fsm.ImplicitSource(lambda x: x ** 2)).integrate(
            u, mesh=[(0, 1, 32), (0, 1, 32)], dt=0.1, step=1)
    with pytest.raises(ValueError):
        fsm.VorticityConvection().integrate(u, mesh=[(0, 1, 32), (0, 1, 32)], dt=0.1, step=1)


@pytest.mark.parametrize("rate", [1.0, 0.5])
def test_ns2d_other_dealiasing_rates_vs_oracle(rate):
    """Without de-aliasing the Nyquist lines take part (self-mirrored lines of the Z-line path)."""
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    n0, n1 = 32, 16
    mesh_info = [(0, 2 * np.pi, n0), (0, 2 * np.pi, n1)]
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(2, 1, n0, n1, generator=g, dtype=torch.float64)
    terms = [("vorticity_convection", -1, {}), ("laplacian", 1 / 50, {})]
    ora = OracleOperator(terms, de_aliasing_rate=rate).register_mesh(mesh_info, 1, dtype="float64")
    ora.set_integrator("ETDRK2")
    want = ora.integrate(u0.numpy(), dt=0.01, step=3)
    op = product_operator(terms)
    op.set_de_aliasing_rate(rate)
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    got = op.integrate(u0, mesh=mesh_info, dt=0.01, step=3)
    assert rel_l2(got.numpy(), want) <= 1e-12


def test_linear_solve_matches_closed_form():
    """LinearOperator.solve (operator/_base.py:217-262): (nu lap - 1) x = b has x_hat = b_hat / (-(nu k^2) - 1)."""
    import torchfsm_b200 as fsm
    n, nu = 32, 0.3
    mesh_info = [(0, 2 * np.pi, n), (0, 2 * np.pi, n)]
    ax = torch.arange(n, dtype=torch.float64) * (2 * np.pi / n)
    x, y = torch.meshgrid(ax, ax, indexing="ij")
    b = (torch.sin(3 * x) * torch.cos(2 * y) + 0.5 * torch.cos(x))[None, None]
    op = nu * fsm.Laplacian() - fsm.ImplicitSource()
    got = op.solve(b, mesh=mesh_info)
    want = torch.sin(3 * x) * torch.cos(2 * y) / (-nu * 13 - 1) + 0.5 * torch.cos(x) / (-nu - 1)
    assert float((got[0, 0] - want).abs().max()) < 1e-13


EXPECTED_KINDS = {"burgers1d_64_etdrk1_f64": [0], "burgers1d_64_etdrk2_f64": [1, 6], "burgers1d_64_setdrk1_f64": [0],
                  "burgers1d_64_setdrk2_f64": [1, 6], "burgers1d_64_setdrk3_f64": [3, 5, 6],
                  "c5_ns3d_16_setdrk4_f64": [3, 4, 5, 6], "c3_ns2d_32_etdrk2_f32": [1, 6], "c2_ks2d_32_f32": [3, 4, 5, 6],
                  "ns2d_32_rk4_f64": [-1, -1, -1, -1]}


@pytest.mark.parametrize("name", sorted(EXPECTED_KINDS))
def test_etd_stages_use_compile_time_combine_shapes(name, monkeypatch):
    """Every stage of the ETD integrators must hit a specialised combine; the generic path gives the same numbers."""
    g = load_golden(name)
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cpu")
    uT = op.integrate(u0, mesh=mesh, dt=spec["dt"], step=spec["steps"])
    assert op._state_dict["integrator"].stage_kinds() == EXPECTED_KINDS[name]
    monkeypatch.setenv("FSM_GENERIC_COMBINE", "1")
    op2, mesh2, _ = product_from_golden(g, "cpu")
    uG = op2.integrate(u0, mesh=mesh2, dt=spec["dt"], step=spec["steps"])
    assert all(k == -1 for k in op2._state_dict["integrator"].stage_kinds())
    assert rel_l2(uT.numpy(), uG.numpy()) <= (1e-6 if name.endswith("f32") else 1e-14)


def test_recorders_real_frames_equal_the_spectrum_protocol():
    """AutoRecorder / CPURecorder take physical frames from the C2R pass; a recorder that only speaks the reference
    protocol (full-spectrum frames, ifftn at the end) must see the same trajectory."""
    import torchfsm_b200 as fsm
    from torchfsm_b200.traj_recorder import _TrajRecorder

    class SpectrumOnly(_TrajRecorder):              # reference-style recorder: no record_real
        def __init__(self, control):
            super().__init__(control)
            self.frames = []

        def _record(self, step, frame):
            assert frame.is_complex() and frame.shape[-1] == frame.shape[-2]      # full spectrum (B, C, n, n)
            self.frames.append(frame.clone())

        @property
        def trajectory(self):
            return torch.fft.ifftn(torch.stack(self.frames, dim=1), dim=(-1, -2)).real

    g = load_golden("c3_ns2d_32_etdrk2_f64")
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cpu")
    ctl = fsm.IntervalController(interval=2, start=1)                                # steps 1, 3, 5
    want = op.integrate(u0, mesh=mesh, dt=spec["dt"], step=5, trajectory_recorder=SpectrumOnly(ctl))
    for cls in (fsm.AutoRecorder, fsm.CPURecorder):
        rec = cls(ctl)
        got = op.integrate(u0, dt=spec["dt"], step=5, trajectory_recorder=rec)
        assert rec._real is True and got.shape == want.shape == (u0.shape[0], 3) + tuple(u0.shape[1:])
        assert rel_l2(got.numpy(), want.numpy()) <= 1e-13
    rec = fsm.AutoRecorder(ctl, include_initial_state=False)
    spec_traj = op.integrate(u0, dt=spec["dt"], step=5, trajectory_recorder=rec, return_in_fourier=True)
    assert rec._real is False and spec_traj.is_complex()
    assert rel_l2(torch.fft.ifftn(spec_traj, dim=(-1, -2)).real.numpy(), want.numpy()) <= 1e-13


@pytest.mark.parametrize("shape", [(8, 8), (16, 8), (8, 16), (64, 8), (8, 64), (256, 8), (8, 256), (512, 8), (8, 512),
                                   (1024, 8), (8, 1024), (8, 128, 8), (8, 256, 8), (8, 512, 8), (512, 8, 8), (8, 8, 512)])
def test_every_line_length_and_axis_role(shape):
    """Each supported line length (8..1024) in each role (first axis = IX/FX, middle axis = MID, last axis = PHYS):
    the plain transforms against torch.fft and one Burgers step against the numpy oracle. Guards the per-size FFT
    decompositions and their compile-time index splits, including the sizes no fixture covers."""
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    nd = len(shape)
    mesh_info = [(0.0, 1.0 + 0.5 * i, n) for i, n in enumerate(shape)]
    gen = torch.Generator().manual_seed(sum(shape))
    u0 = torch.randn(2, nd, *shape, generator=gen, dtype=torch.float64)
    dims = tuple(range(2, 2 + nd))
    u_hat_full = torch.fft.fftn(u0, dim=dims)
    for d, n in zip(dims, shape):                          # band-limit so that one step stays smooth
        f = torch.fft.fftfreq(n, 1.0 / n).abs()
        view = [1] * u0.dim()
        view[d] = n
        u_hat_full = u_hat_full * (f <= max(1, n // 4)).to(u_hat_full.dtype).reshape(view)
    u0 = torch.fft.ifftn(u_hat_full, dim=dims).real.contiguous()
    u0 = u0 / u0.abs().max()
    terms = [("laplacian", 0.01, {}), ("convection", -1, {})]
    op = product_operator(terms)
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    mesh = fsm.MeshGrid(mesh_info, device="cpu", dtype=torch.float64)
    dt = 1e-4
    u1 = op.integrate(u0, mesh=mesh, dt=dt, step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u0)
    assert rel_l2(st.half_to_full(u_hat).numpy(), torch.fft.fftn(u0, dim=dims).numpy()) < 1e-13
    assert rel_l2(st.c2r(u_hat).numpy(), u0.numpy()) < 1e-13
    ora = OracleOperator(terms).register_mesh(mesh_info, nd, dtype="float64")
    ora.set_integrator("ETDRK2")
    integ = ora.build_integrator(dt)
    want = ora.mesh.ifft(integ.step(ora.mesh.fft(u0.numpy()))).real
    assert rel_l2(u1.numpy(), want) < 1e-12


def test_operator_to_changes_the_working_precision():
    """Operator.to(dtype=...) (operator/_base.py:792-803) re-registers the mesh: same operator, fp64 then fp32."""
    g = load_golden("c3_ns2d_32_etdrk2_f64")
    spec = g["spec"]
    op, mesh, u0 = product_from_golden(g, "cpu")
    u64 = op.integrate(u0, mesh=mesh, dt=spec["dt"], step=2)
    op.to(dtype=torch.float32)
    u32 = op.integrate(u0.float(), dt=spec["dt"], step=2)
    assert u32.dtype == torch.float32 and rel_l2(u32.double().numpy(), u64.numpy()) < 1e-5
    assert op._state_dict["integrator"].rdtype == torch.float32


@pytest.mark.parametrize("prog,shape,batch", [("ns2d", (512, 16), 3), ("ns2d", (1024, 8), 2), ("ks2d", (512, 8), 5),
                                              ("ns2d", (16, 512), 2), ("ks2d", (8, 1024), 2), ("ns2d", (256, 32), 3)])
def test_zline_programs_with_long_lines_vs_oracle(prog, shape, batch):
    """The Z-line kernels of the 2-D vorticity / KS programs at the line lengths of C3 (persistent inverse-x kernel with
    cp.async staging for >= 512-point x lines, two-CTA kernels below; phase-overlapping last-axis kernel) on thin grids:
    fp32, two steps, several samples so that the persistent loop iterates."""
    import torchfsm_b200 as fsm
    from oracle import OracleOperator
    mesh_info = [(0.0, 2 * np.pi, shape[0]), (0.0, 2 * np.pi * 1.5, shape[1])]
    g = torch.Generator().manual_seed(sum(shape) + batch)
    u0 = torch.randn((batch, 1) + shape, generator=g, dtype=torch.float64)
    u_hat = torch.fft.fftn(u0, dim=(2, 3))
    for d, n in zip((2, 3), shape):
        f = torch.fft.fftfreq(n, 1.0 / n).abs()
        view = [1, 1, 1, 1]
        view[d] = n
        u_hat = u_hat * (f <= max(1, n // 4)).to(u_hat.dtype).reshape(view)
    u0 = torch.fft.ifftn(u_hat, dim=(2, 3)).real
    u0 = (u0 / u0.abs().max()).float().contiguous()
    if prog == "ns2d":
        terms = [("vorticity_convection", -1, {}), ("laplacian", 0.01, {}), ("implicit_unit_source", -0.1, {})]
        integ_name, dt = "ETDRK2", 1e-3
    else:
        terms = [("laplacian", -0.1, {}), ("biharmonic", -0.01, {}), ("ks_convection", -1, {})]
        integ_name, dt = "SETDRK2", 1e-3
    ora = OracleOperator(terms).register_mesh(mesh_info, 1, dtype="float32")
    ora.set_integrator(integ_name)
    integ = ora.build_integrator(dt)
    op = product_operator(terms)
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2 if prog == "ns2d" else fsm.SETDRKIntegrator.SETDRK2)
    mesh = fsm.MeshGrid(mesh_info, device="cpu", dtype=torch.float32)
    m, c = op._pre_check(u0, None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in integ.tables.items()}
    st = op._build_integrator(dt, batch, tables=tabs)
    got = st.c2r(st.step_half(st.r2c(u0), 2)).numpy()
    want = ora.mesh.ifft(integ.step(integ.step(ora.mesh.fft(u0.numpy())))).real
    assert rel_l2(got, want) < 1e-5
