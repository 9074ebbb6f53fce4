#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on ONE GPU (the bench.py line covers C3).
Prints one JSON line per configuration: ms/step, steps/s, algorithmic GB/s vs the measured HBM peak.

    python tools/bench_configs.py [c1] [c2] [c4] [c5] [--steps K]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfsm_b200 as fsm  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(st, u_hat, steps, warmup=2, one_call=False):
    for _ in range(warmup):
        st.step_half(u_hat, steps if one_call else 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if one_call:
        st.step_half(u_hat, steps)
    else:
        for _ in range(steps):
            st.step_half(u_hat, 1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def pass_breakdown(st, u_hat, steps=2):
    if os.environ.get("FSM_NCU_WINDOW"):          # ncu --profile-from-start off: profile exactly one step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        st.step_half(u_hat, 1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    st.profile(True)
    st.step_half(u_hat, steps)
    torch.cuda.synchronize()
    prof = st.profile_read()
    st.profile(False)
    touched = st.touched_bytes()
    # model_gbs: section 8d model bytes (whole fields per pass; may exceed the HBM peak where only kept modes move);
    # touched_gbs / frac_of_peak: the bytes the pass really reads + writes (kept modes only, tables excluded)
    return {k: {"ms_per_step": round(v["ms"] / steps, 3), "launches_per_step": v["launches"] / steps,
                "model_gbs": round(v["algo_bytes_per_step"] * steps / max(v["ms"], 1e-9) / 1e6, 1),
                "touched_gbs": round(touched[k] * steps / max(v["ms"], 1e-9) / 1e6, 1),
                "frac_of_peak": round(touched[k] * steps / max(v["ms"], 1e-9) / 1e6 / peak(), 3)}
            for k, v in prof.items() if v["launches"]}


def report(name, st, ms, extra=None):
    info = st.info()
    gbs = info["algo_bytes_per_step"] / (ms * 1e-3) / 1e9
    line = {"config": name, "ms_per_step": ms, "steps_per_sec": 1e3 / ms, "algo_gb_per_step": info["algo_bytes_per_step"] / 1e9,
            "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak(), "launches_per_step": info["launches_per_step"],
            "chunk": info["chunk"]}
    touched = sum(st.touched_bytes().values())
    if touched:
        line["touched_gb_per_step"] = touched / 1e9
        line["frac_of_measured_hbm_peak_on_touched_bytes"] = touched / (ms * 1e-3) / 1e9 / peak()
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def c1(steps):
    dev = torch.device("cuda")
    mesh = fsm.MeshGrid([(0, 1, 128)], device=dev, dtype=torch.float32)
    x = mesh.bc_mesh_grid()
    u0 = torch.sin(2 * torch.pi * x) + 0.5
    op = fsm.pde.Burgers(0.01)
    op.integrate(u0, mesh=mesh, dt=0.01, step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u0)
    ms = timed(st, u_hat, 200, warmup=2, one_call=True)
    t0 = time.perf_counter()
    out = op.integrate(u0, dt=0.01, step=200)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    report("C1 burgers1d N=128 B=1 SETDRK4 (200 steps in one launch)", st, ms,
           {"us_per_step": ms * 1e3, "integrate_200_steps_wall_ms": wall * 1e3, "finite": bool(torch.isfinite(out).all())})


def c2(steps):
    dev = torch.device("cuda")
    mesh = fsm.MeshGrid([(0, 60, 256)] * 2, device=dev, dtype=torch.float32)
    u0 = torch.randn(256, 1, 256, 256, generator=torch.Generator().manual_seed(0)).to(dev)
    op = fsm.pde.KuramotoSivashinskyHighDim()
    op.integrate(u0, mesh=mesh, dt=0.5, step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u0)
    ms = timed(st, u_hat, steps)
    report("C2 ks2d 256^2 B=256 SETDRK4", st, ms, {"finite": bool(torch.isfinite(u_hat.real).all()),
                                                    "passes": pass_breakdown(st, u_hat)})


def c4(steps):
    dev = torch.device("cuda")
    n = 256
    mesh = fsm.MeshGrid([(0, 1, n)] * 3, device=dev, dtype=torch.float32)
    x, y, z = mesh.bc_mesh_grid()
    u = torch.cat([torch.sin(2 * np.pi * x) * torch.cos(2 * np.pi * y), torch.cos(2 * np.pi * y) * torch.sin(2 * np.pi * z),
                   torch.sin(2 * np.pi * z) * torch.cos(2 * np.pi * x)], dim=1).repeat(8, 1, 1, 1, 1)
    u = u + 0.1 * fsm.field.diffused_noise(mesh, batch_size=8, n_channel=3, generator=torch.Generator().manual_seed(0))
    op = fsm.pde.Burgers(0.01)
    op.integrate(u, mesh=mesh, dt=0.002, step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u)
    del u
    ms = timed(st, u_hat, steps, warmup=1)
    report("C4 burgers3d 256^3 B=8 C=3 SETDRK4", st, ms, {"finite": bool(torch.isfinite(u_hat.real).all()),
                                                           "passes": pass_breakdown(st, u_hat, 1)})


def c5(steps):
    dev = torch.device("cuda")
    n = 512
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 3, device=dev, dtype=torch.float32)
    x, y, z = mesh.bc_mesh_grid()
    u = torch.cat([torch.sin(x) * torch.cos(y) * torch.cos(z), -torch.cos(x) * torch.sin(y) * torch.cos(z),
                   torch.zeros_like(x)], dim=1)
    del x, y, z
    u = u + 0.05 * fsm.field.diffused_noise(mesh, batch_size=1, n_channel=3, generator=torch.Generator().manual_seed(0))
    op = fsm.pde.NavierStokes(Re=1600)
    op.integrate(u, mesh=mesh, dt=0.0025, step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u)
    del u
    ms = timed(st, u_hat, steps, warmup=1)
    report("C5 ns3d 512^3 B=1 C=3 SETDRK4 (single GPU)", st, ms, {"finite": bool(torch.isfinite(u_hat.real).all()),
                                                                   "passes": pass_breakdown(st, u_hat, 1)})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2", "c4", "c5"])
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    for w in a.which:
        {"c1": c1, "c2": c2, "c4": c4, "c5": c5}[w](a.steps)
