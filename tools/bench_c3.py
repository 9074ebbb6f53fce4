#!/usr/bin/env python
"""C3 step time and per-pass split for one or several builds of the library (kernel tuning).
    python tools/bench_c3.py [--libs a.so,b.so] [--steps 40] [--chunk 0] [--parity]
Each library runs in its own process (the binding loads one library per process)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(a):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import torchfsm_b200 as fsm
    dev = torch.device("cuda", 0)
    n, B = a.n, a.batch
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 2, device=dev, dtype=torch.float32)
    _, y = mesh.bc_mesh_grid()
    op = fsm.pde.NavierStokesVorticity(Re=100.0, force=fsm.field.kolm_force(y))
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    if a.chunk:
        op.set_chunk(a.chunk)
    if a.lanes:
        op.set_lanes(a.lanes)
    u0 = fsm.field.diffused_noise(mesh, batch_size=B, generator=torch.Generator().manual_seed(0))
    op.integrate(u0, mesh=mesh, dt=0.01, step=1)
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u0)
    st.step_half(u_hat, 10)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.step_half(u_hat, a.steps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    st.profile(True)
    st.step_half(u_hat, 10)
    torch.cuda.synchronize()
    prof = st.profile_read()
    st.profile(False)
    out = {"lib": os.path.basename(os.environ.get("FSM_B200_LIB", "libfsm_b200.so")), "ms_per_step": round(best, 4),
           "passes_ms": {k: round(v["ms"] / 10, 4) for k, v in prof.items() if v["launches"]},
           "finite": bool(torch.isfinite(u_hat.real).all()), "chunk": st.info()["chunk"], "lanes": a.lanes}
    if a.parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import test_gpu_parity as t
        errs = t._per_step_vs_oracle([(0, 2 * np.pi, 1024)] * 2, t._ns2d_terms(1024, torch.float32), 1, "ETDRK2", 0.01,
                                     t._smooth((1, 1, 1024, 1024), torch.float32, seed=7), 3, 1e-5)
        out["rel_l2_vs_oracle_3_steps"] = [float("%.3g" % e) for e in errs]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", default="")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child or not a.libs:
        one(a)
    else:
        for lib in a.libs.split(","):
            env = dict(os.environ)
            if lib and lib != "default":
                env["FSM_B200_LIB"] = lib if os.path.isabs(lib) else os.path.join(ROOT, lib)
            args = [sys.executable, os.path.abspath(__file__), "--child", "--steps", str(a.steps), "--chunk", str(a.chunk), "--lanes", str(a.lanes),
                    "--n", str(a.n), "--batch", str(a.batch)] + (["--parity"] if a.parity else [])
            r = subprocess.run(args, env=env, capture_output=True, text=True)
            sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"lib": lib, "error": r.stderr[-400:]}) + "\n")
            sys.stdout.flush()
