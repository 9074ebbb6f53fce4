#!/usr/bin/env python
"""Golden vectors for the operators around the hot path (SURVEY.md §8(f) and the gaps VERDICT r1 lists): the
channel-changing cores (Grad, Div, Curl), the NS diagnostics (Vorticity2Velocity, Vorticity2Pressure,
Velocity2Pressure), ConservativeConvection, ImplicitSource(func), NSPressureConvection with an external force, the
2-D velocity form, tensor-valued (per-sample) coefficients, a linear operator with an explicit source and `solve`.

Runs ONLY in the authoring container: imports the UNMODIFIED reference from /root/reference on CPU and stores
inputs and outputs as tests/golden_ops/<case>_<dtype>.npz. Operators are described by the same term lists
tests/ops_util.py turns into torchfsm_b200 operators.

    python tests/golden/make_golden_ops.py [case ...]
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfsm  # noqa: E402
from torchfsm.mesh import MeshGrid, FourierMesh  # noqa: E402
import torchfsm.operator as ref_ops  # noqa: E402
from torchfsm.integrator import ETDRKIntegrator, SETDRKIntegrator, RKIntegrator  # noqa: E402
from ops_util import build_operator, OPS_CASES, smooth_field, case_sources  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "golden_ops")
torch.set_num_threads(4)
INTEGRATORS = {"ETDRK0": ETDRKIntegrator.ETDRK0, "ETDRK1": ETDRKIntegrator.ETDRK1, "ETDRK2": ETDRKIntegrator.ETDRK2,
               "SETDRK1": SETDRKIntegrator.SETDRK1, "SETDRK2": SETDRKIntegrator.SETDRK2,
               "SETDRK3": SETDRKIntegrator.SETDRK3, "SETDRK4": SETDRKIntegrator.SETDRK4}
INTEGRATORS.update({m.name: m for m in RKIntegrator})


def run_case(case, dtype):
    mesh_info = [tuple(m) for m in case["mesh"]]
    mesh = MeshGrid(mesh_info, dtype=dtype)
    u0 = smooth_field(case, dtype)
    srcs = case_sources(case, dtype)
    out = {"u0": u0.numpy()}
    if case["mode"] == "call":
        op = build_operator(ref_ops, case["terms"], srcs, dtype)
        out["y"] = op(u0.clone(), mesh=mesh).numpy()
    elif case["mode"] == "run_operators":
        ops = [build_operator(ref_ops, t, srcs, dtype) for t in case["operators"]]
        for i, y in enumerate(ref_ops.run_operators(u0.clone(), ops, mesh)):
            out[f"y{i}"] = y.numpy()
    elif case["mode"] == "solve":
        op = build_operator(ref_ops, case["terms"], srcs, dtype)
        out["y"] = op.solve(b=u0.clone(), mesh=mesh, n_channel=case["C"]).numpy()
    else:
        op = build_operator(ref_ops, case["terms"], srcs, dtype)
        if case["integrator"] != "auto":
            op.set_integrator(INTEGRATORS[case["integrator"]])
        out["u1"] = op.integrate(u0.clone(), mesh=mesh, dt=case["dt"], step=1).numpy()
        out["uT"] = op.integrate(u0.clone(), dt=case["dt"], step=case["steps"]).numpy()
        out["rhs0"] = op(u0.clone()).numpy()
    spec = dict(case)
    spec["dtype"] = str(dtype).replace("torch.", "")
    spec["torch"] = torch.__version__
    spec["reference"] = "qiauil/torchfsm v" + getattr(torchfsm, "__version__", "0.0.4")
    out["spec"] = np.array(json.dumps(spec))
    return out


def main():
    only = sys.argv[1:]
    os.makedirs(OUT, exist_ok=True)
    for case in OPS_CASES:
        if only and case["name"] not in only:
            continue
        for dtype in (torch.float32, torch.float64):
            if str(dtype).replace("torch.", "") not in case.get("dtypes", ["float32", "float64"]):
                continue
            out = run_case(case, dtype)
            tag = "f32" if dtype == torch.float32 else "f64"
            path = os.path.join(OUT, f"{case['name']}_{tag}.npz")
            np.savez_compressed(path, **out)
            keys = [k for k in out if k not in ("u0", "spec")]
            fin = all(np.isfinite(out[k]).all() for k in keys)
            print(f"{case['name']:36s} {tag} finite={fin} |out|max={max(np.abs(out[k]).max() for k in keys):.4g} "
                  f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
