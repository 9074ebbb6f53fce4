#!/usr/bin/env python
"""Run ONE timed step of the C3 workload inside a cudaProfilerStart/Stop window, for ncu:
   ncu -f --set full --clock-control none --import-source on --profile-from-start off \
       -o gpurun_out/prof python tools/profile_c3.py [--chunk N] [--batch B]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfsm_b200 as fsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
mesh = fsm.MeshGrid([(0, 2 * np.pi, a.n)] * 2, device=dev, dtype=torch.float32)
_, y = mesh.bc_mesh_grid()
op = fsm.pde.NavierStokesVorticity(Re=100.0, force=fsm.field.kolm_force(y))
op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
if a.chunk:
    op.set_chunk(a.chunk)
u0 = fsm.field.diffused_noise(mesh, batch_size=a.batch, generator=torch.Generator().manual_seed(0))
op.integrate(u0, mesh=mesh, dt=0.01, step=1)
st = op._state_dict["integrator"]
u_hat = st.r2c(u0)
st.step_half(u_hat, 2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.step_half(u_hat, a.steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(u_hat.abs().max()))
