#!/usr/bin/env python
"""C5: 3-D incompressible Navier-Stokes 512^3 (single field, 3 channels, SETDRK4) with the grid
slab-decomposed over the GPUs of one node. Launch with torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29512 \
      tools/bench_slab.py --grid 512 --steps 5

Rank 0 prints one JSON line: ms/step (CUDA events, max over ranks), all-to-all bytes per GPU per step and
the bandwidth the bare exchanges reach, plus a small NCCL parity check against a golden fixture.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torchfsm_b200 as fsm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--integrator", default="SETDRK4")
    ap.add_argument("--nsub", type=int, default=0, help="sub-slabs for exchange/compute overlap (0 = default)")
    ap.add_argument("--modes", default="nccl", help="comma list of exchange[:nsub[:copy_streams[:graph]]] to time; "
                                                    "exchange = nccl, dma or store")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="also time the local passes alone (exchanges skipped) and "
                                                             "per pass class (fsm_profile_*), eager phase loop")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = None
        if os.environ.get("FSM_NCCL_HIGH_PRIORITY", "1") == "1":
            # the local passes occupy every SM; a high-priority NCCL stream lets the exchange's CTAs in
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)

    for mode in a.modes.split(","):
        try:
            mode = mode.strip()
            a.copy_streams, a.graph = 1, 1
            if ":" in mode:                      # "dma:4:2:1" = exchange path : sub-slabs : copy streams : CUDA graph
                f = mode.split(":")
                mode, a.nsub = f[0], int(f[1])
                a.copy_streams = int(f[2]) if len(f) > 2 else 1
                a.graph = int(f[3]) if len(f) > 3 else 1
            run(a, mode, world, rank, dev)
        except Exception as exc:   # report and go on to the next mode
            if rank == 0:
                print(json.dumps({"mode": mode, "n_gpus": world, "error": repr(exc)[:300]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run(a, mode, world, rank, dev):
    # ---- parity of the slab path on a golden fixture (16^3, fp32, SETDRK4)
    parity = None
    if world > 1 and not a.skip_parity:
        from golden_util import load_golden, rel_l2
        from product_util import product_from_golden
        g = load_golden("c5_ns3d_16_setdrk4_f32")
        if g["u0"].shape[2] % world == 0 and g["u0"].shape[2] // world >= 2:
            op, mesh, u0 = product_from_golden(g, dev)
            op.set_slab_decomposition(exchange=mode)
            nxl = u0.shape[2] // world
            sl = slice(rank * nxl, (rank + 1) * nxl)
            uT = op.integrate(u0[:, :, sl].contiguous(), mesh=mesh, dt=g["spec"]["dt"], step=g["spec"]["steps"])
            err = torch.tensor([rel_l2(uT.cpu().numpy(), g["uT"][:, :, sl])], device=dev, dtype=torch.float64)
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            parity = float(err.item())

    # ---- the big grid
    n = a.grid
    mesh = fsm.MeshGrid([(0, 2 * np.pi, n)] * 3, device=dev, dtype=torch.float32)
    nxl = n // world
    ax = torch.arange(n, device=dev, dtype=torch.float32) * (2 * np.pi / n)
    x = ax[rank * nxl:(rank + 1) * nxl].reshape(1, 1, nxl, 1, 1)
    y = ax.reshape(1, 1, 1, n, 1)
    z = ax.reshape(1, 1, 1, 1, n)
    u = torch.cat([torch.sin(x) * torch.cos(y) * torch.cos(z), -torch.cos(x) * torch.sin(y) * torch.cos(z),
                   0.05 * torch.sin(3 * y + x) * torch.cos(2 * z)], dim=1).contiguous()
    op = fsm.pde.NavierStokes(Re=1600)
    op.set_integrator(getattr(fsm.SETDRKIntegrator, a.integrator) if a.integrator.startswith("S")
                      else getattr(fsm.ETDRKIntegrator, a.integrator))
    if world > 1:
        op.set_slab_decomposition(nsub=a.nsub, exchange=mode, copy_streams=a.copy_streams, graph=bool(a.graph))
    t0 = time.perf_counter()
    op.integrate(u, mesh=mesh, dt=0.0025, step=1)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    st = op._state_dict["integrator"]
    u_hat = st.r2c(u)
    del u

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        st.step_half(u_hat, 1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.step_half(u_hat, a.steps)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    finite = bool(torch.isfinite(u_hat.real).all())

    breakdown = None
    if a.breakdown and world > 1:
        # (1) the local passes alone: same phase loop, exchanges replaced by nothing (results are garbage, times are not)
        saved_exchange, saved_graph = st._exchange, st._graph_enabled
        st._graph_enabled = False
        st._exchange = lambda which, count, offset=0, async_op=False: (fsm.operator._StreamWork(st.device) if async_op else None)
        st.step_half(u_hat.clone(), 1)
        barrier()
        e0.record()
        st.step_half(u_hat.clone(), a.steps)
        e1.record()
        barrier()
        t_local = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        # (2) per pass class, exchanges still skipped (serialised events)
        st.profile(True)
        st.step_half(u_hat.clone(), 2)
        torch.cuda.synchronize()
        prof = st.profile_read()
        st.profile(False)
        st._exchange = saved_exchange
        # (3) eager phase loop with the real exchanges (no CUDA graph)
        st.step_half(u_hat.clone(), 1)
        barrier()
        e0.record()
        st.step_half(u_hat.clone(), a.steps)
        e1.record()
        barrier()
        t_eager = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t_eager, op=dist.ReduceOp.MAX)
        st._graph_enabled = saved_graph
        breakdown = {"local_only_ms_per_step": float(t_local.item()), "eager_ms_per_step": float(t_eager.item()),
                     "passes_ms_per_step": {k: round(v["ms"] / 2, 3) for k, v in prof.items()},
                     "pass_launches_per_step": {k: v["launches"] // 2 for k, v in prof.items()}}

    a2a = None
    if world > 1:
        c1, c2 = st._slab_counts[0]
        n_eval = st.n_stages
        bytes_per_step = n_eval * (c1 + c2) * 8 * (world - 1) / world      # leaves this GPU per step
        barrier()
        e0.record()
        for _ in range(3 * n_eval):
            st._exchange(0, c1)
            st._exchange(1, c2)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        a2a = {"send_bytes_per_gpu_per_step": bytes_per_step, "bare_exchange_ms_per_step": float(t.item()),
               "bare_exchange_gbs_per_gpu": bytes_per_step / (float(t.item()) * 1e-3) / 1e9}
        if mode == "store":   # the "exchange" is only the barrier pair: the bytes moved inside the kernels
            a2a = {"send_bytes_per_gpu_per_step": bytes_per_step, "barriers_ms_per_step": float(t.item())}
    if rank == 0:
        info = st.info()
        print(json.dumps({"config": f"C5 ns3d {n}^3 B=1 C=3 {a.integrator} slab x{world}", "mode": mode, "n_gpus": world,
                          "ms_per_step": float(ms.item()), "steps_per_sec": 1e3 / float(ms.item()), "steps": a.steps,
                          "finite": finite, "setup_s": setup_s, "nsub": getattr(st, "nsub", 1), "copy_streams": a.copy_streams,
                          "graph": bool(getattr(st, "_graphs", None)), "graph_error": getattr(st, "_graph_error", None), "all_to_all": a2a, "breakdown": breakdown, "parity_rel_l2_16cubed_fp32": parity,
                          "algo_gb_per_step_global": info["algo_bytes_per_step"] / 1e9}), flush=True)
    del st, op, u_hat
    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
