"""Time integrators of the oracle (numpy). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Restates:
  ETDRK0/1/2 tables and steps      integrator/_etdrk.py:10-82   (Cox & Matthews 2002)
  SETDRK1-4 contour-mean tables    integrator/_stable_etdrk/_uncached.py:8-211 (Kassam & Trefethen 2005)
  SETDRK1-4 step formulas          integrator/_stable_etdrk/_setdrk_step.py:5-82
  fixed-step RK4                   integrator/_rk.py:43-58, 142-155
Every integrator exposes ``.dt``, ``.step(u_hat)`` and ``.tables`` (dict of arrays) so
tests can feed the very same coefficient tables to the CUDA path (the
plain-ETDRK tables cancel catastrophically in fp32, so "same tables" is part of
"same inputs").
"""

import numpy as np


def _cdt(L):
    return np.complex64 if L.dtype in (np.complex64, np.float32) else np.complex128


def _rdt(L):
    return np.float32 if L.dtype in (np.complex64, np.float32) else np.float64


class ETDRK0:
    order_name = "ETDRK0"

    def __init__(self, dt, L, nonlinear_func=None, tables=None):
        self.dt = dt
        self.N = nonlinear_func
        if tables is not None:
            self.tables = dict(tables)
        else:
            self.tables = self._build(dt, L)

    @staticmethod
    def _build(dt, L):
        rd = _rdt(L)
        return {"exp": np.exp(rd(dt) * L)}                              # _etdrk.py:21

    def step(self, u):
        return self.tables["exp"] * u                                    # _etdrk.py:27


class ETDRK1(ETDRK0):
    order_name = "ETDRK1"

    @staticmethod
    def _build(dt, L):
        t = ETDRK0._build(dt, L)
        rd, cd = _rdt(L), _cdt(L)
        Ls = np.where(L == 0, cd(1), L)
        t["coef_1"] = np.where(L == 0, cd(dt), (t["exp"] - rd(1)) / Ls).astype(t["exp"].dtype)   # _etdrk.py:43-45
        return t

    def step(self, u):
        t = self.tables
        return t["exp"] * u + t["coef_1"] * self.N(u)                    # _etdrk.py:51


class ETDRK2(ETDRK1):
    order_name = "ETDRK2"

    @staticmethod
    def _build(dt, L):
        t = ETDRK1._build(dt, L)
        rd, cd = _rdt(L), _cdt(L)
        Ls = np.where(L == 0, cd(1), L)
        t["coef_2"] = np.where(
            L == 0, cd(dt / 2), (t["exp"] - rd(1) - L * rd(dt)) / (Ls * Ls * rd(dt))
        ).astype(t["exp"].dtype)                                         # _etdrk.py:66-70
        return t

    def step(self, u):
        t = self.tables
        n0 = self.N(u)                                                   # _etdrk.py:76-82
        a = t["exp"] * u + t["coef_1"] * n0
        n1 = self.N(a)
        return a + t["coef_2"] * (n1 - n0)


# ------------------------------------------------------------------ stable ETDRK (contour)
def _roots_of_unity(M, rd):
    # _uncached.py:8-15: exp(2 pi i (j - 1/2) / M), j = 1..M, built in the real dtype
    j = np.arange(1, M + 1).astype(rd)
    cd = np.complex64 if rd == np.float32 else np.complex128
    return np.exp(cd(2j * np.pi) * (j - rd(0.5)) / rd(M)).astype(cd)


def _lr(dt, L, M, radius):
    rd = _rdt(L)
    return rd(radius) * _roots_of_unity(M, rd) + L[..., None] * rd(dt)   # _uncached.py:34-43


def _cmean(x, rd):
    # torch.mean(..., axis=-1).real
    return x.mean(axis=-1).real.astype(rd)


class _SETDRKBase:
    def __init__(self, dt, L, nonlinear_func, n_integration_points=16, integration_radius=1.0, tables=None):
        self.dt = dt
        self.N = nonlinear_func
        if tables is not None:
            self.tables = dict(tables)
        else:
            self.tables = self._build(dt, L, n_integration_points, integration_radius)


class SETDRK1(_SETDRKBase):
    order_name = "SETDRK1"

    @staticmethod
    def _build(dt, L, M, radius):
        rd = _rdt(L)
        lr = _lr(dt, L, M, radius)
        return {"exp": np.exp(rd(dt) * L),
                "coef_1": rd(dt) * _cmean((np.exp(lr) - 1) / lr, rd)}      # _uncached.py:72

    def step(self, u):
        t = self.tables
        return t["exp"] * u + t["coef_1"] * self.N(u)                    # _setdrk_step.py:5-11


class SETDRK2(_SETDRKBase):
    order_name = "SETDRK2"

    @staticmethod
    def _build(dt, L, M, radius):
        rd = _rdt(L)
        lr = _lr(dt, L, M, radius)
        return {"exp": np.exp(rd(dt) * L),
                "coef_1": rd(dt) * _cmean((np.exp(lr) - 1) / lr, rd),             # _uncached.py:103
                "coef_2": rd(dt) * _cmean((np.exp(lr) - 1 - lr) / lr ** 2, rd)}   # _uncached.py:104

    def step(self, u):
        t = self.tables
        n0 = self.N(u)                                                   # _setdrk_step.py:14-25
        a = t["exp"] * u + t["coef_1"] * n0
        n1 = self.N(a)
        return a + t["coef_2"] * (n1 - n0)


class SETDRK3(_SETDRKBase):
    order_name = "SETDRK3"

    @staticmethod
    def _build(dt, L, M, radius):
        rd = _rdt(L)
        lr = _lr(dt, L, M, radius)
        e = np.exp(lr)
        return {
            "exp": np.exp(rd(dt) * L),
            "half_exp": np.exp(rd(0.5) * rd(dt) * L),                                            # _uncached.py:135
            "coef_1": rd(dt) * _cmean((np.exp(lr / 2) - 1) / lr, rd),                            # :136
            "coef_2": rd(dt) * _cmean((e - 1) / lr, rd),                                         # :137
            "coef_3": rd(dt) * _cmean((-4 - lr + e * (4 - 3 * lr + lr ** 2)) / (lr ** 3), rd),   # :138-143
            "coef_4": rd(dt) * _cmean((4.0 * (2.0 + lr + e * (-2 + lr))) / (lr ** 3), rd),       # :144-149
            "coef_5": rd(dt) * _cmean((-4 - 3 * lr - lr ** 2 + e * (4 - lr)) / (lr ** 3), rd),   # :150-156
        }

    def step(self, u):
        t = self.tables
        n0 = self.N(u)                                                   # _setdrk_step.py:28-52
        a = t["half_exp"] * u + t["coef_1"] * n0
        n1 = self.N(a)
        b = t["exp"] * u + t["coef_2"] * (2 * n1 - n0)
        n2 = self.N(b)
        return t["exp"] * u + t["coef_3"] * n0 + t["coef_4"] * n1 + t["coef_5"] * n2


class SETDRK4(_SETDRKBase):
    order_name = "SETDRK4"

    @staticmethod
    def _build(dt, L, M, radius):
        rd = _rdt(L)
        lr = _lr(dt, L, M, radius)
        e = np.exp(lr)
        c1 = rd(dt) * _cmean((np.exp(lr / 2) - 1) / lr, rd)                                      # _uncached.py:192
        return {
            "exp": np.exp(rd(dt) * L),
            "half_exp": np.exp(rd(0.5) * rd(dt) * L),                                            # :191
            "coef_1": c1, "coef_2": c1, "coef_3": c1,                                            # :193-195
            "coef_4": rd(dt) * _cmean((-4 - lr + e * (4 - 3 * lr + lr ** 2)) / (lr ** 3), rd),   # :196-201
            "coef_5": rd(dt) * _cmean((2 + lr + e * (-2 + lr)) / (lr ** 3), rd),                 # :202-205
            "coef_6": rd(dt) * _cmean((-4 - 3 * lr - lr ** 2 + e * (4 - lr)) / (lr ** 3), rd),   # :206-211
        }

    def step(self, u):
        t = self.tables
        n0 = self.N(u)                                                   # _setdrk_step.py:55-82
        a = t["half_exp"] * u + t["coef_1"] * n0
        n1 = self.N(a)
        b = t["half_exp"] * u + t["coef_2"] * n1
        n2 = self.N(b)
        c = t["half_exp"] * a + t["coef_3"] * (2 * n2 - n0)
        n3 = self.N(c)
        return t["exp"] * u + t["coef_4"] * n0 + t["coef_5"] * 2 * (n1 + n2) + t["coef_6"] * n3


# ------------------------------------------------------------------------- classical RK4
class RK4:
    """Fixed-step RK4 on f(u_hat) = L u_hat + N(u_hat) (_rk.py:43-58, 142-155; _base.py:408-439)."""
    order_name = "RK4"
    ca = [[1 / 2, 1 / 2], [1 / 2, 0, 1 / 2], [1, 0, 0, 1]]
    b = [1 / 6, 1 / 3, 1 / 3, 1 / 6]

    def __init__(self, dt, rhs):
        self.dt = dt
        self.f = rhs
        self.tables = {}

    def step(self, x):
        dt = self.dt
        ks = [self.f(x)]
        for row in self.ca:
            ks.append(self.f(x + dt * sum(a * k for a, k in zip(row[1:], ks))))
        return x + dt * sum(b * k for b, k in zip(self.b, ks))


# the other fixed-step members of the explicit Runge-Kutta family: rows [c_i, a_i1, ...] and weights b as the reference
# holds them (_rk.py:82-255), stepped by the same loop (_rk.py:43-58)
RK_FAMILY = {
    "Euler": ([[1.0]], [1.0]),                                                                       # _rk.py:82-87
    "Midpoint": ([[1 / 2, 1 / 2]], [0, 1]),                                                          # :90-95
    "Heun12": ([[1, 1]], [1 / 2, 1 / 2]),                                                            # :98-105
    "Ralston12": ([[2 / 3, 2 / 3]], [1 / 4, 3 / 4]),                                                 # :108-120
    "BogackiShampine23": ([[1 / 2, 1 / 2], [3 / 4, 0, 3 / 4], [1, 2 / 9, 1 / 3, 4 / 9]], [2 / 9, 1 / 3, 4 / 9, 0]),   # :123-139
    "RK4_38Rule": ([[1 / 3, 1 / 3], [2 / 3, -1 / 3, 1], [1, -1, 1, 1]], [1 / 8, 3 / 8, 3 / 8, 1 / 8]),               # :158-171
    "Dorpi45": ([[1 / 5, 1 / 5], [3 / 10, 3 / 40, 9 / 40], [4 / 5, 44 / 45, -56 / 15, 32 / 9],      # :174-202
                 [8 / 9, 19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
                 [1, 9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
                 [1, 35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]],
                [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]),
    "Fehlberg45": ([[1 / 4, 1 / 4], [3 / 8, 3 / 32, 9 / 32], [12 / 13, 1932 / 2197, -7200 / 2197, 7296 / 2197],     # :205-225
                    [1, 439 / 216, -8, 3680 / 513, -845 / 4104], [1 / 2, -8 / 27, 2, -3544 / 2565, 1859 / 4104, -11 / 40]],
                   [16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55]),
    "CashKarp45": ([[1 / 5, 1 / 5], [3 / 10, 3 / 40, 9 / 40], [3 / 5, 3 / 10, -9 / 10, 6 / 5],      # :228-255
                    [1, -11 / 54, 5 / 2, -70 / 27, 35 / 27],
                    [7 / 8, 1631 / 55296, 175 / 512, 575 / 13824, 44275 / 110592, 253 / 4096]],
                   [37 / 378, 0, 250 / 621, 125 / 594, 0, 512 / 1771]),
}


def explicit_rk(name, dt, rhs):
    """An RK4-style stepper with the tableau of ``name``."""
    r = RK4(dt, rhs)
    r.order_name = name
    r.ca, r.b = RK_FAMILY[name]
    return r


INTEGRATORS = {c.order_name: c for c in (ETDRK0, ETDRK1, ETDRK2, SETDRK1, SETDRK2, SETDRK3, SETDRK4)}
