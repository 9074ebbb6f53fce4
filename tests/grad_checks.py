"""Shared body of the gradient-mode checks (emulator on CPU, CUDA library on the GPU)."""
import torch

from grad_util import grad_case, load, namespace, run

TOL = {torch.float64: 1e-10, torch.float32: 2e-4}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def check_grad_case(name, device, dtype):
    """integrate(u0) and op(u0) with u0.requires_grad through torchfsm_b200: values and gradients against the vectors
    the reference produced under its own autograd (tests/golden_grad)."""
    case, gold = grad_case(name), load(name, dtype)
    got = run(namespace("b200"), case, dtype, device)
    assert set(got) == set(gold) - {"spec"}
    for k in sorted(got):
        assert got[k].shape == gold[k].shape, k
        assert torch.isfinite(got[k]).all(), k
        err = rel(got[k].cpu(), gold[k])
        tol = case.get("tol32", {}).get(k, TOL[dtype]) if dtype == torch.float32 else TOL[dtype]
        assert err < tol, f"{name} {k}: {err:.3e}"
