"""Integrator selection and coefficient tables (host-side set-up, torch ops on the target device).

Mirrors the reference's enums (``ETDRKIntegrator``, ``SETDRKIntegrator``, ``RKIntegrator``) and
evaluates the same closed-form expressions for the tables so the CUDA path consumes the very
numbers the reference would:
  ETDRK0/1/2   integrator/_etdrk.py:21, 43-45, 66-70        (Cox & Matthews 2002)
  SETDRK1-4    integrator/_stable_etdrk/_uncached.py:8-211  (Kassam & Trefethen 2005 contour means)
The contour tables are evaluated in slabs, so the 16x temporary of the reference
(``_uncached.py:34-43``; OOM at 512^3) never materialises.
"""
from enum import Enum

import torch


class ETDRKIntegrator(Enum):
    ETDRK0 = "ETDRK0"
    ETDRK1 = "ETDRK1"
    ETDRK2 = "ETDRK2"


class SETDRKIntegrator(Enum):
    SETDRK1 = "SETDRK1"
    SETDRK2 = "SETDRK2"
    SETDRK3 = "SETDRK3"
    SETDRK4 = "SETDRK4"


class RKIntegrator(Enum):
    """integrator/_rk.py:257-282. RK4 runs as fused stages inside the library; the other members run as right-hand-side
    evaluations (``fsm_rhs``) combined by ``fsm_lincomb`` (non-adaptive stepping, the reference's default)."""
    Euler = "Euler"
    Midpoint = "Midpoint"
    Heun12 = "Heun12"
    Ralston12 = "Ralston12"
    BogackiShampine23 = "BogackiShampine23"
    RK4 = "RK4"
    RK4_38Rule = "RK4_38Rule"
    Dorpi45 = "Dorpi45"
    Fehlberg45 = "Fehlberg45"
    CashKarp45 = "CashKarp45"


# Butcher tableaus in the reference's row format [c_i, a_i1, a_i2, ...] (integrator/_rk.py:82-255; textbook constants) and
# the weights b. _RKBase._rk_step evaluates k_0 = f(x), then one k per row, and x' = x + dt * sum(b_i k_i) over
# zip(b, ks): a row may be evaluated and not used (Euler's, Dormand-Prince's last).
RK_TABLEAUS = {
    "Euler": ([[1.0]], [1.0]),
    "Midpoint": ([[1 / 2, 1 / 2]], [0, 1]),
    "Heun12": ([[1, 1]], [1 / 2, 1 / 2]),
    "Ralston12": ([[2 / 3, 2 / 3]], [1 / 4, 3 / 4]),
    "BogackiShampine23": ([[1 / 2, 1 / 2], [3 / 4, 0, 3 / 4], [1, 2 / 9, 1 / 3, 4 / 9]], [2 / 9, 1 / 3, 4 / 9, 0]),
    "RK4_38Rule": ([[1 / 3, 1 / 3], [2 / 3, -1 / 3, 1], [1, -1, 1, 1]], [1 / 8, 3 / 8, 3 / 8, 1 / 8]),
    "Dorpi45": ([[1 / 5, 1 / 5], [3 / 10, 3 / 40, 9 / 40], [4 / 5, 44 / 45, -56 / 15, 32 / 9],
                 [8 / 9, 19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
                 [1, 9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
                 [1, 35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
                [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]),
    "Fehlberg45": ([[1 / 4, 1 / 4], [3 / 8, 3 / 32, 9 / 32], [12 / 13, 1932 / 2197, -7200 / 2197, 7296 / 2197],
                    [1, 439 / 216, -8, 3680 / 513, -845 / 4104], [1 / 2, -8 / 27, 2, -3544 / 2565, 1859 / 4104, -11 / 40]],
                   [16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55]),
    "CashKarp45": ([[1 / 5, 1 / 5], [3 / 10, 3 / 40, 9 / 40], [3 / 5, 3 / 10, -9 / 10, 6 / 5],
                    [1, -11 / 54, 5 / 2, -70 / 27, 35 / 27],
                    [7 / 8, 1631 / 55296, 175 / 512, 575 / 13824, 44275 / 110592, 253 / 4096]],
                   [37 / 378, 0, 250 / 621, 125 / 594, 0, 512 / 1771]),
}


def has_imag(t) -> bool:
    """Does a table / symbol carry a genuine imaginary part? (2j*pi*f)**4 and **6 carry ~1e-13 of rounding in the
    imaginary part although the symbol is real (generic/_spatial_derivative.py:7-20 accepts them): compare against the
    magnitude, not against 0."""
    if t is None or not t.is_complex():
        return False
    t = t.detach()
    eps = torch.finfo(t.real.dtype).eps
    return float(t.imag.abs().max()) > 8 * eps * float(t.abs().max())


def integrator_name(integrator, is_linear):
    if isinstance(integrator, str):
        if integrator != "auto":
            raise AssertionError("The integrator should be 'auto' or an instance of ETDRKIntegrator, "
                                 "SETDRKIntegrator or RKIntegrator")
        return "ETDRK0" if is_linear else "SETDRK4"       # operator/_base.py:451-455
    if isinstance(integrator, (ETDRKIntegrator, SETDRKIntegrator, RKIntegrator)):
        return integrator.value
    raise AssertionError("The integrator should be 'auto' or an instance of ETDRKIntegrator, "
                         "SETDRKIntegrator or RKIntegrator")


def etdrk_tables(name, dt, L):
    """Plain ETD tables; ``L`` complex, any shape."""
    t = {"exp": torch.exp(dt * L)}
    if name in ("ETDRK1", "ETDRK2"):
        t["coef_1"] = torch.where(L == 0, dt, (t["exp"] - 1) / L)
    if name == "ETDRK2":
        t["coef_2"] = torch.where(L == 0, dt / 2, (t["exp"] - 1 - L * dt) / (L ** 2 * dt))
    return t


def _roots_of_unity(M, device, dtype):
    return torch.exp(2j * torch.pi * (torch.arange(1, M + 1, device=device, dtype=dtype) - 0.5) / M)


def setdrk_tables(name, dt, L, n_integration_points=16, integration_radius=1.0, slab=1 << 20):
    """Contour-integral tables of the stable ETDRK schemes, evaluated slab by slab."""
    t = {"exp": torch.exp(dt * L)}
    if name in ("SETDRK3", "SETDRK4"):
        t["half_exp"] = torch.exp(0.5 * dt * L)
    roots = integration_radius * _roots_of_unity(n_integration_points, L.device, L.real.dtype)
    flat = L.reshape(-1)
    names = {"SETDRK1": ["coef_1"], "SETDRK2": ["coef_1", "coef_2"],
             "SETDRK3": ["coef_1", "coef_2", "coef_3", "coef_4", "coef_5"],
             "SETDRK4": ["coef_1", "coef_4", "coef_5", "coef_6"]}[name]
    out = {k: torch.empty(flat.shape, device=L.device, dtype=L.real.dtype) for k in names}
    for s in range(0, flat.numel(), slab):
        lr = roots + flat[s:s + slab].unsqueeze(-1) * dt
        sl = slice(s, s + slab)

        def mean(x):
            return dt * torch.mean(x, axis=-1).real

        if name == "SETDRK1":
            out["coef_1"][sl] = mean((torch.exp(lr) - 1) / lr)
        elif name == "SETDRK2":
            out["coef_1"][sl] = mean((torch.exp(lr) - 1) / lr)
            out["coef_2"][sl] = mean((torch.exp(lr) - 1 - lr) / lr ** 2)
        elif name == "SETDRK3":
            e = torch.exp(lr)
            out["coef_1"][sl] = mean((torch.exp(lr / 2) - 1) / lr)
            out["coef_2"][sl] = mean((e - 1) / lr)
            out["coef_3"][sl] = mean((-4 - lr + e * (4 - 3 * lr + lr ** 2)) / (lr ** 3))
            out["coef_4"][sl] = mean((4.0 * (2.0 + lr + e * (-2 + lr))) / (lr ** 3))
            out["coef_5"][sl] = mean((-4 - 3 * lr - lr ** 2 + e * (4 - lr)) / (lr ** 3))
        else:
            e = torch.exp(lr)
            out["coef_1"][sl] = mean((torch.exp(lr / 2) - 1) / lr)
            out["coef_4"][sl] = mean((-4 - lr + e * (4 - 3 * lr + lr ** 2)) / (lr ** 3))
            out["coef_5"][sl] = mean((2 + lr + e * (-2 + lr)) / (lr ** 3))
            out["coef_6"][sl] = mean((-4 - 3 * lr - lr ** 2 + e * (4 - lr)) / (lr ** 3))
    for k, v in out.items():
        t[k] = v.reshape(L.shape)
    if name == "SETDRK4":                                   # _uncached.py:193-195
        t["coef_2"] = t["coef_1"]
        t["coef_3"] = t["coef_1"]
    return t


def build_tables(name, dt, L, **cfg):
    if name.startswith("SETDRK"):
        cfg = {k: v for k, v in cfg.items() if k in ("n_integration_points", "integration_radius")}
        return setdrk_tables(name, dt, L, **cfg)
    if name.startswith("ETDRK"):
        return etdrk_tables(name, dt, L)
    return {}
