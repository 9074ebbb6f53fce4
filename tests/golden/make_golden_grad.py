#!/usr/bin/env python
"""Golden gradients for torchfsm_b200's gradient mode (torchfsm_b200/autograd.py): the UNMODIFIED reference from
/root/reference integrates / evaluates the operators of tests/grad_util.py on CPU under torch autograd, and the values
and the gradients with respect to the initial field are stored as tests/golden_grad/<case>_<dtype>.npz.

Runs ONLY in the authoring container.      python tests/golden/make_golden_grad.py [case ...]
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
from grad_util import GRAD_CASES, GRAD_DIR, STEPS, namespace, run  # noqa: E402

torch.set_num_threads(4)


def main():
    only = sys.argv[1:]
    os.makedirs(GRAD_DIR, exist_ok=True)
    ns = namespace("reference")
    for case in GRAD_CASES:
        if only and case["name"] not in only:
            continue
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            out = {k: v.numpy() for k, v in run(ns, case, dtype).items()}
            spec = dict(case, dtype=tag, steps=STEPS, torch=torch.__version__,
                        reference="qiauil/torchfsm v" + getattr(torchfsm, "__version__", "0.0.4"))
            out["spec"] = np.array(json.dumps(spec))
            path = os.path.join(GRAD_DIR, f"{case['name']}_{tag}.npz")
            np.savez_compressed(path, **out)
            fin = all(np.isfinite(v).all() for k, v in out.items() if k != "spec")
            print(f"{case['name']:34s} {tag} finite={fin} |grad_y|max={np.abs(out['grad_y']).max():.4g} "
                  f"{os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
