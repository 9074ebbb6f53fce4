#!/usr/bin/env python
"""Local passes of the slab-decomposed C5 step on ONE GPU, no communication: a plan for rank 0 of P ranks is stepped with
the exchanges replaced by nothing (the receive buffers keep whatever they hold: times are real, values are not).
Separates the cost of the rank-blocked / cyclic addressing and of the smaller launches from the cost of the exchange.

    python tools/slab_local_profile.py [--grid 512] [--ranks 1,2,4,8] [--steps 3]
    ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/slab_local_profile.py --ranks 2 --steps 1
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfsm_b200 as fsm  # noqa: E402
from torchfsm_b200.operator import _StreamWork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=512)
ap.add_argument("--ranks", default="1,2,4,8")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--nsub", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
n = a.grid
for P in [int(x) for x in a.ranks.split(",")]:
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 3, device=dev, dtype=torch.float32)
    op = fsm.pde.NavierStokes(Re=1600)
    op.set_integrator(fsm.SETDRKIntegrator.SETDRK4)
    if P > 1:
        op.set_slab_decomposition(group=None, rank=0, nranks=P, nsub=a.nsub, exchange="nccl", graph=False)
    nxl = n // P
    u = torch.randn(1, 3, nxl, n, n, device=dev) * 0.01
    m, c = op._pre_check(u, None, mesh)
    op.register_mesh(m, c)
    st = op._build_integrator(0.0025, 1)
    if P > 1:
        st._exchange = lambda which, count, offset=0, async_op=False: (_StreamWork(st.device) if async_op else None)
        for t in st._recv + st._send:
            t.zero_()
    u_hat = st.empty_half().zero_()
    st.step_half(u_hat, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.step_half(u_hat, a.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    st.profile(True)
    st.step_half(u_hat, 2)
    torch.cuda.synchronize()
    prof = st.profile_read()
    st.profile(False)
    print(json.dumps({"ranks": P, "grid": n, "local_ms_per_step": round(ms, 3), "ideal_ms": None,
                      "passes_ms_per_step": {k: round(v["ms"] / 2, 3) for k, v in prof.items()},
                      "launches_per_step": {k: v["launches"] // 2 for k, v in prof.items()}}), flush=True)
    del st, op, u_hat, u
    torch.cuda.empty_cache()
