#!/usr/bin/env python
"""Benchmark of the hot path: 2-D Navier-Stokes (vorticity form, Kolmogorov forcing) 1024^2, batch 64
per GPU, ETDRK2 with 2/3 dealiasing (BASELINE.json configs[2] = the configuration the metric is quoted on).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "spectral step" advances ONE 1024^2 sample by ONE ETDRK2 time step; `value` is the whole-job rate
in sample-steps/s with the state resident in HBM (rot-half spectra), `e2e` the same metric through the
public API (`Operator.integrate`) with host buffers and both PCIe copies inside the timed region.
Multi-GPU: the ensemble shards over ranks with no data-path collective (weak scaling, 64 samples/GPU).
`--impl reference` times the CPU oracle port of the reference's path (the reference is pure Python over
torch and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID, BATCH, DT, RE = 1024, 64, 0.01, 100.0
WORKLOAD = "ns2d_vorticity_kolmogorov_1024x1024_b64_etdrk2_dealias23_fp32"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_operator(workers):
    from oracle import OracleOperator
    ax = np.arange(N_GRID, dtype=np.float32) * np.float32(2 * np.pi / N_GRID)
    y = np.broadcast_to(ax.reshape(1, 1, 1, N_GRID), (1, 1, N_GRID, N_GRID))
    src = (4.0 * np.cos(4.0 * y)).astype(np.float32)
    op = OracleOperator([("vorticity_convection", -1, {}), ("laplacian", 1 / RE, {}),
                         ("implicit_unit_source", -0.1, {}), ("explicit_source", -1, {"source": src})])
    op.register_mesh([(0, 2 * np.pi, N_GRID)] * 2, 1, dtype="float32", workers=workers)
    op.set_integrator("ETDRK2")
    return op


def time_oracle(sample_batch, steps, warmup):
    """CPU port of the reference path (oracle/) on the host cores; returns sample-steps/s."""
    cores = os.cpu_count() or 1
    op = oracle_operator(cores)
    integ = op.build_integrator(DT)
    rng = np.random.default_rng(0)
    u = rng.standard_normal((sample_batch, 1, N_GRID, N_GRID)).astype(np.float32)
    u_hat = op.mesh.fft(u) * op.mesh.low_pass_filter(0.1)
    for _ in range(warmup):
        u_hat = integ.step(u_hat)
    t0 = time.perf_counter()
    for _ in range(steps):
        u_hat = integ.step(u_hat)
    dt = time.perf_counter() - t0
    assert np.isfinite(u_hat).all()
    return sample_batch * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 4
    value, ms, cores = time_oracle(sample_batch, args.steps, args.warmup)
    sample = (f"{sample_batch} of {BATCH} samples per step at the full 1024^2 grid, {args.steps} timed steps; "
              "rate is per sample so no extrapolation is involved")
    line = {
        "impl": "reference", "metric": "spectral_steps_per_sec", "value": value, "unit": "sample-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "grid": [N_GRID, N_GRID], "batch_per_gpu": BATCH, "integrator": "ETDRK2",
                   "dt": DT, "note": "CPU oracle port of the reference path (numpy + scipy.fft, all host threads)"},
        "cpu_baseline": {"value": value, "unit": "sample-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sample-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torchfsm_b200 as fsm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    mesh = fsm.MeshGrid([(0, 2 * np.pi, N_GRID)] * 2, device=dev, dtype=torch.float32)
    _, y = mesh.bc_mesh_grid()
    op = fsm.pde.NavierStokesVorticity(Re=RE, force=fsm.field.kolm_force(y))
    op.set_integrator(fsm.ETDRKIntegrator.ETDRK2)
    if args.chunk:
        op.set_chunk(args.chunk)
    gen = torch.Generator().manual_seed(rank)          # every rank owns a different ensemble shard
    u0 = fsm.field.diffused_noise(mesh, batch_size=BATCH, generator=gen)
    # registers the mesh and builds the stepper once; later calls omit `mesh` (reference quirk Q3)
    op.integrate(u0, mesh=mesh, dt=DT, step=1)
    st = op._state_dict["integrator"]
    info = st.info()
    u_hat = st.r2c(u0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput
    # one fsm_step call advances all K steps (the FX of a stage is fused with the IX of the next one,
    # also across step boundaries); the state (269 MB) and scratch exceed the L2, so no flush is needed
    st.step_half(u_hat, args.warmup)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.step_half(u_hat, args.steps)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    assert torch.isfinite(u_hat.real).all(), "state blew up"
    ms_step = ms_total / args.steps
    value = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---------------- per-pass timing for the roofline entry (same state, right after the timed region)
    st.profile(True)
    prof_steps = max(3, min(10, args.steps))
    st.step_half(u_hat, prof_steps)
    torch.cuda.synchronize()
    prof = st.profile_read()
    st.profile(False)
    peak, peak_src = _peaks()
    dom = max((k for k in prof if prof[k]["launches"] > 0), key=lambda k: prof[k]["ms"])
    passes = {}
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        gbs = v["algo_bytes_per_step"] * prof_steps / (v["ms"] * 1e-3) / 1e9
        passes[k] = {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                     "algo_gb_per_step": v["algo_bytes_per_step"] / 1e9, "achieved_gbs": gbs, "frac": gbs / peak}
    # DRAM traffic per launch of the dominant kernel from the committed ncu --set full capture (same chunk)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        if info["chunk"] == 64:
            traffic = tj["kernels"]["k_pass_" + dom.lower()]["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    step_gbs = info["algo_bytes_per_step"] / (ms_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_pass_" + dom.lower(), "achieved": passes[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": passes[dom]["frac"], "traffic": traffic,
                "traffic_source": "profiles/r1_ncu_traffic.json (ncu --set full, dram read+write bytes per launch)",
                "algo_bytes_per_launch": passes[dom]["algo_gb_per_step"] * 1e9 / passes[dom]["launches_per_step"],
                "peak_source": peak_src,
                "passes": passes,
                "whole_step": {"algo_gb_per_step": info["algo_bytes_per_step"] / 1e9, "achieved_gbs": step_gbs,
                               "frac": step_gbs / peak}}

    # ---------------- end to end through the public API with host buffers
    host_in = torch.empty((BATCH, 1, N_GRID, N_GRID), dtype=torch.float32, pin_memory=True)
    host_in.copy_(u0)
    host_out = torch.empty_like(host_in, pin_memory=True)
    dev_in = torch.empty_like(u0)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        dev_in.copy_(host_in, non_blocking=True)
        host_out.copy_(op.integrate(dev_in, dt=DT, step=1), non_blocking=True)
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        dev_in.copy_(host_in, non_blocking=True)
        out = op.integrate(dev_in, dt=DT, step=1)
        host_out.copy_(out, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * BATCH * e2e_steps / (ms_e2e * 1e-3)
    nbytes = host_in.numel() * 4

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sb = 4
            v, ms, cores = time_oracle(sb, 4, 1)
            cpu = {"value": v, "unit": "sample-steps/s", "cores": cores, "kind": "port",
                   "sample": f"{sb} of {BATCH} samples at the full 1024^2 grid, 4 timed ETDRK2 steps after 1 warm-up "
                             f"({ms:.0f} ms per {sb}-sample step); numpy+scipy.fft oracle port, all host threads"}
        line = {
            "metric": "spectral_steps_per_sec", "value": value, "unit": "sample-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "grid": [N_GRID, N_GRID], "batch_per_gpu": BATCH, "integrator": "ETDRK2",
                       "dt": DT, "Re": RE, "dealias": "2/3", "chunk": info["chunk"],
                       "batch_steps_per_sec": 1e3 / ms_step,
                       "l2_policy": "state + scratch arrays (3 x 269 MB) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"ensemble x{world} (no collective)"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "sample-steps/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "api": "Operator.integrate(u_0, dt, step=1) incl. r2c + ETDRK2 step + c2r, pinned host buffers"},
            "gpu_launches": int(info["launches_per_step"] * args.steps),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunk", type=int, default=0, help="samples per pass launch (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
