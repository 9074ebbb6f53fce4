#!/usr/bin/env python
"""MEASURED COMPARISON ONLY (never the product path): the reference's own GPU path for C3, restated
op-for-op in torch — full C2C complex64 spectra through torch.fft (cuFFT) plus un-fused ATen
element-wise kernels, exactly the sequence `Operator.integrate` runs per ETDRK2 step
(operator/_base.py:375-403, dedicated/_navier_stokes.py:41-46, _base.py:1007-1015,
integrator/_etdrk.py:72-82): 12 ifftn + 2 fftn per step, two of the six inverse transforms per
evaluation wasted. Prints one JSON line with ms/step on this GPU.

    python tools/torch_cufft_comparison.py [--batch 64] [--steps 10]
"""
import argparse
import json

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev, n, B, dt, Re = torch.device("cuda"), a.n, a.batch, 0.01, 100.0
f = torch.fft.fftfreq(n, 2 * np.pi / n, device=dev, dtype=torch.float32)
bfx, bfy = f.reshape(1, 1, n, 1), f.reshape(1, 1, 1, n)
gx, gy = (2j * torch.pi * bfx) ** 1, (2j * torch.pi * bfy) ** 1
lap = (2j * torch.pi * bfx) ** 2 + (2j * torch.pi * bfy) ** 2
inv_lap = torch.where(lap == 0, 1.0, 1 / lap)
mask = torch.ones_like(lap.real)
for bf in (bfx, bfy):
    mask = mask * torch.where(bf.abs() > bf.abs().max() * (2 / 3), 0, 1)
L = (1 / Re) * lap + (-0.1) * torch.ones_like(bfx)
E = torch.exp(dt * L)
c1 = torch.where(L == 0, dt, (E - 1) / L)
c2 = torch.where(L == 0, dt / 2, (E - 1 - L * dt) / (L ** 2 * dt))
y = (torch.arange(n, device=dev, dtype=torch.float32) * (2 * np.pi / n)).reshape(1, 1, 1, n).expand(1, 1, n, n)
src_hat = torch.fft.fftn(4.0 * torch.cos(4.0 * y), dim=(-2, -1))
dims = (-1, -2)


def nonlinear(u_fft):
    result = 0.0
    d_fft = u_fft * mask
    _ = torch.fft.ifftn(d_fft, dim=dims).real                     # shared dealiased_u, unused by the core
    psi = -d_fft * inv_lap
    ux = torch.fft.ifftn(gy * psi, dim=dims).real
    uy = torch.fft.ifftn(-gx * psi, dim=dims).real
    wx = torch.fft.ifftn(gx * d_fft, dim=dims).real
    wy = torch.fft.ifftn(gy * d_fft, dim=dims).real
    result = result + (-1) * torch.fft.fftn(ux * wx + uy * wy, dim=dims)
    _ = torch.fft.ifftn(u_fft, dim=dims).real                     # u for the explicit source, unused
    result = result + (-1) * src_hat
    return result


def step(u):
    n0 = nonlinear(u)
    s1 = E * u + c1 * n0
    n1 = nonlinear(s1)
    return s1 + c2 * (n1 - n0)


g = torch.Generator().manual_seed(0)
u0 = torch.randn(B, 1, n, n, generator=g).to(dev)
u = torch.fft.fftn(u0, dim=dims) * torch.exp(lap.real)            # diffused noise
u = u / u.abs().amax()
for _ in range(2):
    u = step(u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    u = step(u)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"what": "reference op sequence in torch (cuFFT C2C + ATen), measured comparison only",
                  "config": f"NS2D vorticity + Kolmogorov forcing {n}^2 x {B}, ETDRK2, fp32", "ms_per_step": ms,
                  "sample_steps_per_sec": B * 1e3 / ms, "finite": bool(torch.isfinite(u.real).all()),
                  "torch": torch.__version__}))
