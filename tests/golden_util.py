"""Helpers shared by the parity tests: load golden fixtures, build oracle operators."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(dtype_tag=None):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if dtype_tag:
        names = [n for n in names if n.endswith("_" + dtype_tag)]
    return names


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["spec"] = json.loads(str(g["spec"]))
    return g


def spec_terms(g):
    """Term list of a fixture with explicit-source arrays substituted for their recipe names."""
    terms = []
    for i, (kind, coef, params) in enumerate(g["spec"]["terms"]):
        params = dict(params)
        if kind == "explicit_source":
            params["source"] = g[f"src_{i}"]
        terms.append((kind, coef, params))
    return terms


def oracle_from_golden(g, workers=1):
    from oracle import OracleOperator
    spec = g["spec"]
    op = OracleOperator(spec_terms(g))
    op.register_mesh([tuple(m) for m in spec["mesh"]], spec["C"], dtype=spec["dtype"], workers=workers)
    op.set_integrator(spec["integrator"])
    return op


def golden_tables(g):
    names = {"tab_exp_term": "exp", "tab_half_exp_term": "half_exp"}
    names.update({f"tab_coef_{i}": f"coef_{i}" for i in range(1, 7)})
    return {names[k]: v for k, v in g.items() if k in names}


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))
