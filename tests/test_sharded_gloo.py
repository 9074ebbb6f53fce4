"""Multi-process path on CPU (gloo, world_size 2): the ensemble shards over ranks with no data-path
collective; every rank integrates its shard through the same plugin layer + C ABI (host-emulator build)
and the gathered result equals the single-process reference answer."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, name, lib_path, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from golden_util import load_golden, rel_l2
    from product_util import product_from_golden
    from torchfsm_b200 import _cabi
    _cabi.use_library(lib_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        spec = g["spec"]
        op, mesh, u0 = product_from_golden(g, "cpu")
        per = u0.shape[0] // world
        shard = u0[rank * per:(rank + 1) * per].contiguous()
        uT = op.integrate(shard, mesh=mesh, dt=spec["dt"], step=spec["steps"])
        gathered = [torch.empty_like(uT) for _ in range(world)]
        dist.all_gather(gathered, uT)
        # device-style timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            full = torch.cat(gathered, dim=0)
            out_q.put((rel_l2(full.numpy(), g["uT"]), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c3_ns2d_32_etdrk2_f64", "c4_burgers3d_16_f64"])
def test_ensemble_shards_over_two_ranks(name):
    from product_util import build_emulator
    lib_path = build_emulator()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, lib_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    err, tmax = q.get(timeout=5)
    assert err <= 1e-12 and tmax == 2.0


def _slab_worker(rank, world, port, name, lib_path, out_q, nsub=0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from golden_util import load_golden, rel_l2
    from product_util import product_from_golden
    from torchfsm_b200 import _cabi
    _cabi.use_library(lib_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        spec = g["spec"]
        op, mesh, u0 = product_from_golden(g, "cpu")
        if nsub in ("store", "dma"):   # peer-visible receive buffers: kernels (store) or block copies (dma) fill them
            from product_util import SharedFilePeers
            op.set_slab_decomposition(exchange=nsub, peers=SharedFilePeers(rank, world, os.environ["FSM_TEST_PEER_DIR"]))
        else:
            op.set_slab_decomposition(nsub=nsub)
        nxl = u0.shape[2] // world
        local = u0[:, :, rank * nxl:(rank + 1) * nxl].contiguous()       # this rank's physical x-slab
        uT = op.integrate(local, mesh=mesh, dt=spec["dt"], step=spec["steps"])
        rhs = op(local)
        st = op._state_dict["integrator"]
        back = st.c2r(st.r2c(local))
        errs = torch.tensor([rel_l2(uT.numpy(), g["uT"][:, :, rank * nxl:(rank + 1) * nxl]),
                             rel_l2(rhs.numpy(), g["rhs0"][:, :, rank * nxl:(rank + 1) * nxl]),
                             rel_l2(back.numpy(), local.numpy())], dtype=torch.float64)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        if rank == 0:
            out_q.put(errs.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world,nsub", [("c5_ns3d_16_setdrk4_f64", 2, 1), ("c4_burgers3d_16_f64", 2, 2),
                                             ("c5_ns3d_32x16x8_etdrk2_f64", 2, 4), ("burgers3d_8x16x32_rk4_f64", 2, 0),
                                             ("c5_ns3d_32x16x8_etdrk2_f32", 4, 2),
                                             ("c5_ns3d_16_setdrk4_f64", 2, "store"), ("c4_burgers3d_16_f64", 4, "store"),
                                             ("c5_ns3d_32x16x8_etdrk2_f32", 4, "store"),
                                             ("c5_ns3d_16_setdrk4_f64", 2, "dma"), ("c5_ns3d_32x16x8_etdrk2_f32", 4, "dma")])
def test_slab_decomposed_grid_matches_reference(name, world, nsub, tmp_path, monkeypatch):
    """ONE 3-D grid split into x-slabs (physical) / ky-slabs (spectral) over the ranks; the transposes are
    all_to_all_single (gloo here, NCCL on GPUs); results equal the reference's single-device answer."""
    from product_util import build_emulator
    lib_path = build_emulator()
    monkeypatch.setenv("FSM_TEST_PEER_DIR", str(tmp_path))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, name, lib_path, q, nsub)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    e_step, e_rhs, e_rt = q.get(timeout=5)
    tol = 1e-12 if name.endswith("f64") else 1e-5
    assert e_step <= tol * (1 if name.endswith("f64") else 3) and e_rhs <= 10 * tol and e_rt <= tol


def _ks_worker(rank, world, port, name, lib_path, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from golden_util import load_golden, rel_l2
    from product_util import product_from_golden
    from torchfsm_b200 import _cabi
    _cabi.use_library(lib_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(name)
        spec = g["spec"]
        op, mesh, u0 = product_from_golden(g, "cpu")
        op.set_ensemble_group(dist.group.WORLD)
        cut = u0.shape[0] - 1                                          # uneven shards when B = 3: 2 + 1 samples
        lo, hi = (0, cut) if rank == 0 else (cut, u0.shape[0])
        uT = op.integrate(u0[lo:hi].contiguous(), mesh=mesh, dt=spec["dt"], step=spec["steps"])
        err = torch.tensor([rel_l2(uT.numpy(), g["uT"][lo:hi])], dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            out_q.put(float(err.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c2_ks2d_32_f64", "ks3d_16_f64"])
def test_ks_ensemble_mean_spans_ranks(name):
    """KS batch mean across ranks: local means + one all-reduce of the zero-mode log."""
    from product_util import build_emulator
    lib_path = build_emulator()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ks_worker, args=(r, 2, port, name, lib_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) <= 1e-12


def _slab_ops_worker(rank, world, port, names, lib_path, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torchfsm_b200 as fsm
    from golden_util import rel_l2
    from ops_util import build_operator, case_sources, load_ops
    from torchfsm_b200 import _cabi
    _cabi.use_library(lib_path)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        worst = 0.0
        for name in names:
            g = load_ops(name)
            spec = g["spec"]
            dtype = torch.float64
            mesh = fsm.MeshGrid([tuple(m) for m in spec["mesh"]], dtype=dtype)
            op = build_operator(fsm, spec["terms"], case_sources(spec, dtype), dtype)
            op.set_slab_decomposition()
            u0 = torch.from_numpy(g["u0"])
            nxl = u0.shape[2] // world
            y = op(u0[:, :, rank * nxl:(rank + 1) * nxl].contiguous(), mesh=mesh)
            worst = max(worst, rel_l2(y.numpy(), g["y"][:, :, rank * nxl:(rank + 1) * nxl]))
        err = torch.tensor([worst], dtype=torch.float64)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        if rank == 0:
            out_q.put(float(err.item()))
    finally:
        dist.destroy_process_group()


def test_spectral_maps_and_pressure_on_a_slab_decomposed_grid():
    """Grad / Div / Curl / Velocity2Pressure of ONE 3-D grid whose x-slabs live on two ranks: the point-wise maps act on
    the local ky lines (cyclic ownership), the transforms around them exchange through all_to_all_single."""
    from product_util import build_emulator
    lib_path = build_emulator()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    names = ["grad3d_f64", "div3d_f64", "curl3d_f64", "vel2p_3d_f64"]
    procs = [ctx.Process(target=_slab_ops_worker, args=(r, 2, port, names, lib_path, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) <= 1e-11
