#!/usr/bin/env python
"""Cost of the two unrolled modes (torchfsm_b200/autograd.py, unrolled.py) next to the fused step and to the unmodified
reference on the same GPU: 2-D Navier-Stokes vorticity, ETDRK2, fp32.

  fused      : nu*Lap - VorticityConvection                       (the fused kernels)
  paired     : the same + beta * d/dx (complex symbol)            (pair of half spectra, unrolled)
  gradient   : forward + backward of a quadratic loss after STEPS steps (gradient mode)
  reference  : the same three through baseline/_ref (cuFFT + ATen), when it is importable

    python tools/bench_unrolled.py [--grid 1024] [--batch 16] [--steps 8]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchfsm_b200 as fsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=1024)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
n, B, S = a.grid, a.batch, a.steps


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def suite(ns, mesh_cls, etd):
    mesh = mesh_cls([(0, 2 * np.pi, n)] * 2, device=dev, dtype=torch.float32)
    g = torch.Generator().manual_seed(0)
    u0 = (0.5 * torch.randn(B, 1, n, n, generator=g)).to(dev)
    u0 = (0.002 * ns.Laplacian()).integrate(u0, mesh=mesh, dt=1.0, step=1)
    plain = 0.01 * ns.Laplacian() - ns.VorticityConvection()
    beta = 0.01 * ns.Laplacian() - ns.VorticityConvection() + 0.5 * ns.SpatialDerivative(0, 1)
    for op in (plain, beta):
        op.set_integrator(etd.ETDRK2)
    out = {}
    out["fused_ms_per_step"] = timed(lambda: plain.integrate(u0, mesh=mesh, dt=0.005, step=S)) / S
    out["paired_ms_per_step"] = timed(lambda: beta.integrate(u0, mesh=mesh, dt=0.005, step=S)) / S

    def fwd_bwd():
        x = u0.clone().requires_grad_(True)
        (plain.integrate(x, mesh=mesh, dt=0.005, step=S) ** 2).sum().backward()
        return x.grad
    out["gradient_fwd_bwd_ms_per_step"] = timed(fwd_bwd) / S
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    return out


res = {"grid": n, "batch": B, "steps": S, "dtype": "f32", "b200": suite(fsm, fsm.MeshGrid, fsm.ETDRKIntegrator)}
torch.cuda.reset_peak_memory_stats()
try:
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import torchfsm.operator as rop
    import torchfsm.integrator as rint
    from torchfsm.mesh import MeshGrid as RefMesh
    res["reference"] = suite(rop, RefMesh, rint.ETDRKIntegrator)
except Exception as e:                                  # noqa: BLE001
    res["reference"] = {"unavailable": repr(e)[:200]}
print(json.dumps(res))
