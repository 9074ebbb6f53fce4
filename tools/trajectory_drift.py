#!/usr/bin/env python
"""Trajectory drift over a full run (north star: "trajectory drift over the full run reported").

Runs the product (CUDA library on a GPU box, otherwise the host-emulator build of the same kernel sources)
and the numpy oracle side by side from the same initial condition and the same coefficient tables and
prints, per config, the relative L2 distance of the physical fields after 1, 10, 50, 100, ... steps up to the
full run. Chaotic configs (KS, NS) amplify rounding differences: reported, not gated.

    python tools/trajectory_drift.py [c1] [c3] [c2] [c2full] [c4] [c5]     # (bounded versions of) the BASELINE configs
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torchfsm_b200 as fsm  # noqa: E402
from torchfsm_b200 import _cabi  # noqa: E402
from oracle import OracleOperator  # noqa: E402


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def backend():
    if torch.cuda.is_available() and os.path.exists(_cabi.DEFAULT_LIB):
        return "cuda", torch.device("cuda", 0)
    from product_util import build_emulator
    _cabi.use_library(build_emulator())
    return "emulator (same kernel sources on the CPU)", torch.device("cpu")


def drift(name, mesh_info, terms_fn, integrator, dt, steps, u0, dtype, dev, marks):
    from product_util import product_operator
    nd = len(mesh_info)
    np_dtype = "float32" if dtype == torch.float32 else "float64"
    mesh = fsm.MeshGrid(mesh_info, device=dev, dtype=dtype)
    terms_t = terms_fn(lambda a: torch.from_numpy(a).to(dev))
    terms_n = terms_fn(lambda a: a)
    ora = OracleOperator(terms_n).register_mesh(mesh_info, u0.shape[1], dtype=np_dtype, workers=os.cpu_count())
    ora.set_integrator(integrator)
    integ = ora.build_integrator(dt)
    op = product_operator(terms_t)
    enum = fsm.SETDRKIntegrator if integrator.startswith("S") else fsm.ETDRKIntegrator
    op.set_integrator(getattr(enum, integrator))
    ud = torch.from_numpy(u0).to(dev)
    m, c = op._pre_check(ud, None, mesh)
    op.register_mesh(m, c)
    tabs = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in integ.tables.items()}
    st = op._build_integrator(dt, u0.shape[0], tables=tabs)
    u_hat = st.r2c(ud)
    ref_hat = ora.mesh.fft(u0)
    out, done = {}, 0
    for mark in [m_ for m_ in marks if m_ <= steps] + ([steps] if steps not in marks else []):
        st.step_half(u_hat, mark - done)
        for _ in range(mark - done):
            ref_hat = integ.step(ref_hat)
        done = mark
        out[str(mark)] = rel_l2(st.c2r(u_hat).cpu().numpy(), ora.mesh.ifft(ref_hat).real)
    return {"config": name, "dtype": np_dtype, "integrator": integrator, "dt": dt, "steps": steps,
            "grid": [m_[2] for m_ in mesh_info], "batch": int(u0.shape[0]), "rel_l2_after_steps": out}


def main():
    which = sys.argv[1:] or ["c1", "c3", "c2"]
    name, dev = backend()
    marks = [1, 10, 50, 100, 200]
    res = {"backend": name, "note": "same tables on both sides; oracle = numpy restatement of the reference", "runs": []}
    for dtype in (torch.float32, torch.float64):
        npd = np.float32 if dtype == torch.float32 else np.float64
        if "c1" in which:   # C1 exactly: 1-D Burgers nu=0.01, N=128, dt=0.01, 200 steps, SETDRK4
            x = (np.arange(128) / 128.0).astype(npd).reshape(1, 1, 128)
            u0 = (np.sin(2 * np.pi * x) + 0.5).astype(npd)
            res["runs"].append(drift("C1 burgers1d (full config)", [(0.0, 1.0, 128)],
                                     lambda cv: [("laplacian", 0.01, {}), ("convection", -1, {})], "SETDRK4", 0.01, 200,
                                     u0, dtype, dev, marks))
        if "c3" in which:   # C3 bounded: 2-D NS vorticity + Kolmogorov forcing, 256^2 x 2 (full: 1024^2 x 64), ETDRK2
            n = 256
            ax = (np.arange(n) * (2 * np.pi / n)).astype(npd)
            src = (4.0 * np.cos(4.0 * ax)).reshape(1, 1, 1, n).repeat(n, axis=2).astype(npd)
            g = torch.Generator().manual_seed(0)
            meshc = fsm.MeshGrid([(0, 2 * np.pi, n)] * 2, device=dev, dtype=dtype)
            u0 = fsm.field.diffused_noise(meshc, batch_size=2, generator=g).cpu().numpy().astype(npd)
            res["runs"].append(drift("C3 ns2d vorticity, Kolmogorov forcing (256^2 x 2 of 1024^2 x 64)",
                                     [(0, 2 * np.pi, n)] * 2,
                                     lambda cv: [("vorticity_convection", -1, {}), ("laplacian", 1 / 100, {}),
                                                 ("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": cv(src)})],
                                     "ETDRK2", 0.01, 200, u0, dtype, dev, marks))
        if "c2" in which:   # C2 bounded: 2-D KS 64^2 x 4 (full: 256^2 x 256), SETDRK4, 200 steps
            n = 64
            g = np.random.default_rng(0)
            u0 = g.standard_normal((4, 1, n, n)).astype(npd)
            res["runs"].append(drift("C2 ks2d (64^2 x 4 of 256^2 x 256)", [(0, 60.0 * n / 256, n)] * 2,
                                     lambda cv: [("laplacian", -1, {}), ("biharmonic", -1, {}),
                                                 ("ks_convection", -1, {"remove_mean": True})],
                                     "SETDRK4", 0.1, 200, u0, dtype, dev, marks))
        if "c2full" in which:   # C2 at its own grid and time step: 2-D KS 256^2 (L = 60), batch bounded to 8, SETDRK4, dt = 0.5
            n = 256
            g = np.random.default_rng(0)
            u0 = g.standard_normal((8, 1, n, n)).astype(npd)
            res["runs"].append(drift("C2 ks2d (256^2 x 8 of 256^2 x 256)", [(0, 60.0, n)] * 2,
                                     lambda cv: [("laplacian", -1, {}), ("biharmonic", -1, {}),
                                                 ("ks_convection", -1, {"remove_mean": True})],
                                     "SETDRK4", 0.5, 200, u0, dtype, dev, marks))
        if "c4" in which:   # C4 bounded: 3-D Burgers 64^3 x 2 (full: 256^3 x 8), SETDRK4, dt = 0.002, 50 steps
            n = 64
            ax = (np.arange(n) / n).astype(npd) * 2 * np.pi
            x, y, z = ax.reshape(1, 1, n, 1, 1), ax.reshape(1, 1, 1, n, 1), ax.reshape(1, 1, 1, 1, n)
            base = np.concatenate([np.sin(x) * np.cos(y) * np.cos(z), -np.cos(x) * np.sin(y) * np.cos(z),
                                   0.5 * np.sin(2 * x + z) * np.cos(y)], axis=1).astype(npd)
            g = np.random.default_rng(1)
            u0 = (np.repeat(base, 2, axis=0) + 0.05 * np.sin(x + 2 * y) * g.standard_normal((2, 3, 1, 1, 1))).astype(npd)
            res["runs"].append(drift("C4 burgers3d (64^3 x 2 of 256^3 x 8)", [(0, 1.0, n)] * 3,
                                     lambda cv: [("laplacian", 0.01, {}), ("convection", -1, {})], "SETDRK4", 0.002, 50, u0,
                                     dtype, dev, [1, 10, 50]))
        if "c5" in which:   # C5 bounded: 3-D NS 64^3 (full: 512^3), Taylor-Green + perturbation, SETDRK4, dt = 0.0025 * 8, 20 steps
            n = 64
            ax = (np.arange(n) / n).astype(npd) * 2 * np.pi
            x, y, z = ax.reshape(1, 1, n, 1, 1), ax.reshape(1, 1, 1, n, 1), ax.reshape(1, 1, 1, 1, n)
            u0 = np.concatenate([np.sin(x) * np.cos(y) * np.cos(z) + 0.05 * np.sin(2 * y) * np.cos(3 * z) + 0 * x,
                                 -np.cos(x) * np.sin(y) * np.cos(z) + 0.05 * np.sin(3 * z + x) + 0 * y,
                                 0.05 * np.sin(3 * y + x) * np.cos(2 * z)], axis=1).astype(npd)
            res["runs"].append(drift("C5 ns3d velocity form (64^3 of 512^3)", [(0, 2 * np.pi, n)] * 3,
                                     lambda cv: [("ns_pressure_convection", 1, {}), ("laplacian", 1 / 1600, {})], "SETDRK4",
                                     0.02, 20, u0, dtype, dev, [1, 10, 20]))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
