"""Oracle: vector fields of the named systems (TEST INFRASTRUCTURE ONLY).

NumPy restatement of reference systems.py (`_f_np` branches) and utils.py
(Normalize).  The PDE systems keep the reference's DENSE difference matrices and
`@` mat-vecs (systems.py:321-353, 365-366, 421-450) so the oracle rounds like the
reference does.
"""
import numpy as np


class Normalizer:
    """utils.py:1-33 -- identity or affine map to [-1, 1]."""

    def __init__(self, mn, mx, kind=None):
        self.mn = np.asarray(mn, dtype=float)
        self.mx = np.asarray(mx, dtype=float)
        kind = "identity" if kind is None else kind.lower()
        if kind not in ("identity", "-11"):
            raise NotImplementedError("Only identity and -11 are implemented")
        self.kind = kind

    def fit(self, x):
        if self.kind == "-11":
            return 2 * (x - self.mn) / (self.mx - self.mn) - 1
        return x

    def inverse(self, x):
        if self.kind == "-11":
            return (x + 1) / 2 * (self.mx - self.mn) + self.mn
        return x

    def scale(self):
        if self.kind == "-11":
            return 2 / (self.mx - self.mn)
        return 1


class OracleSystem:
    """systems.py:23-77 (`ODE`): name, u0 (normalised), normalised vector field."""

    def __init__(self, name, mn, mx, u0, normalization=None):
        self.name = name
        self.norm = Normalizer(mn, mx, normalization)
        self.u0 = np.array(self.norm.fit(np.asarray(u0, dtype=float)), dtype=float)

    def f_orig(self, t, u):
        raise NotImplementedError

    def f(self, t, u):
        # systems.py:36-40
        v = self.norm.inverse(u)
        out = self.f_orig(t, v)
        return out * self.norm.scale()

    def dim(self):
        return self.u0.shape[0]


class FHN_ODE(OracleSystem):
    def __init__(self, **kw):
        super().__init__("FHN_ODE", [-2, -1], [2.1, 1.2], [-1, 1], **kw)

    def f_orig(self, t, u):
        # systems.py:97-106
        a, b, c = 0.2, 0.2, 3
        out = np.zeros(u.shape)
        out[0] = c * (u[0] - ((u[0] ** 3) / 3) + u[1])
        out[1] = -(1 / c) * (u[0] - a + b * u[1])
        return out


class Rossler(OracleSystem):
    def __init__(self, **kw):
        super().__init__("Rossler", [-10, -11, 0], [12, 8, 23], [0, -6.78, 0.02], **kw)

    def f_orig(self, t, u):
        # systems.py:128-137
        a, b, c = 0.2, 0.2, 5.7
        out = np.zeros(u.shape)
        out[0] = -u[1] - u[2]
        out[1] = u[0] + (a * u[1])
        out[2] = b + u[2] * (u[0] - c)
        return out


class Hopf(OracleSystem):
    def __init__(self, tspan=(-20, 500), **kw):
        self.maxtime = tspan[1]
        super().__init__("Hopf", [-23, -23, 0], [23, 23, 1], [0.1, 0.1, tspan[0]], **kw)

    def f_orig(self, t, u):
        # systems.py:157-163
        mt = self.maxtime
        out = np.zeros(u.shape)
        out[0] = -u[1] + u[0] * ((u[2] / mt) - u[0] ** 2 - u[1] ** 2)
        out[1] = u[0] + u[1] * ((u[2] / mt) - u[0] ** 2 - u[1] ** 2)
        out[2] = 1
        return out


class DblPend(OracleSystem):
    def __init__(self, **kw):
        super().__init__("DblPend", [-2, -2.5, -17, -3.5], [2, 2.5, 1, 3.5], [-0.5, 0, 0, 0], **kw)

    def f_orig(self, t, u):
        # systems.py:191-199
        out = np.zeros(u.shape)
        out[0] = u[1]
        out[1] = (-1 / (2 - np.cos(u[0] - u[2]) ** 2)) * ((u[1] ** 2) * np.cos(u[0] - u[2]) * np.sin(u[0] - u[2]) + (u[3] ** 2) * np.sin(u[0] - u[2]) + 2 * np.sin(u[0]) - np.cos(u[0] - u[2]) * np.sin(u[2]))
        out[2] = u[3]
        out[3] = (-1 / (2 - np.cos(u[0] - u[2]) ** 2)) * (-2 * (u[1] ** 2) * np.sin(u[0] - u[2]) - (u[3] ** 2) * np.sin(u[0] - u[2]) * np.cos(u[0] - u[2]) - 2 * np.cos(u[0] - u[2]) * np.sin(u[0]) + 2 * np.sin(u[2]))
        return out


class Brusselator(OracleSystem):
    def __init__(self, **kw):
        super().__init__("Brusselator", [0.4, 0.9], [4, 5], [1, 3.07], **kw)

    def f_orig(self, t, u):
        # systems.py:217-222
        out = np.zeros(u.shape)
        out[0] = 1 + (u[0] ** 2) * u[1] - (3 + 1) * u[0]
        out[1] = 3 * u[0] - (u[0] ** 2) * u[1]
        return out


class Lorenz(OracleSystem):
    def __init__(self, **kw):
        super().__init__("Lorenz", [-17.1, -23, 6], [18.1, 25, 45], [-15, -15, 20], **kw)

    def f_orig(self, t, u):
        # systems.py:241-247
        out = np.zeros(u.shape)
        out[0] = 10 * (u[1] - u[0])
        out[1] = 28 * u[0] - u[1] - u[0] * u[2]
        out[2] = u[0] * u[1] - (8 / 3) * u[2]
        return out


class ThomasLabyrinth(OracleSystem):
    def __init__(self, **kw):
        super().__init__("ThomasLabyrinth", [-12, -12, -12], [12, 12, 12],
                         [4.6722764, 5.2437205e-10, -6.4444208e-10], **kw)

    def f_orig(self, t, u):
        # systems.py:273-288
        a, b = 0.5, 10.0
        out = np.zeros(u.shape)
        out[0] = -a * u[0] + b * np.sin(u[1])
        out[1] = -a * u[1] + b * np.sin(u[2])
        out[2] = -a * u[2] + b * np.sin(u[0])
        return out


def _periodic_second_diff(n, scale):
    # systems.py:327-343 / 425-437: tridiagonal (1,-2,1) with periodic corners
    T = np.diag(-2 * np.ones(n))
    i = np.arange(n - 1)
    T[i, i + 1] = 1
    T[i + 1, i] = 1
    D = scale * T
    D[0, -1] = scale
    D[-1, 0] = scale
    return D


class FHN_PDE(OracleSystem):
    def __init__(self, d_x, seed=45, **kw):
        # systems.py:292-318; IC = legacy np.random.seed(seed) uniform stream
        self.d_x = d_x
        d = 2 * d_x * d_x
        dx = 2 / (d_x - 1)
        Dxx = _periodic_second_diff(d_x, 1 / (dx ** 2))
        self.DXX = np.kron(np.eye(d_x, d_x), Dxx)
        self.DYY = np.kron(Dxx, np.eye(d_x, d_x))
        np.random.seed(seed)
        rng = np.random.Generator(np.random.get_bit_generator())
        u0 = rng.uniform(size=d)
        super().__init__(f"FHN_PDE_{d_x}", [-1] * d, [1] * d, u0, **kw)

    def f_orig(self, t, u):
        # systems.py:370-383
        d = int(u.shape[0] / 2)
        u1, u2 = u[:d], u[d:]
        a, b, k, tau = 2.8e-4, 5e-3, -5e-3, 0.1
        U = a * (self.DXX + self.DYY) @ u1 + u1 - (u1 ** 3) - u2 + k * np.ones(d)
        V = (1 / tau) * (b * (self.DXX + self.DYY) @ u2 + u1 - u2)
        return np.hstack([U, V])


class Burgers(OracleSystem):
    def __init__(self, d_x, nu=1 / 100, **kw):
        # systems.py:403-442
        self.d_x, self.nu = d_x, nu
        d = d_x
        dx = 2 / (d - 1)
        self.Dxx = _periodic_second_diff(d, nu / (dx ** 2))
        Tx = np.zeros((d, d))
        i = np.arange(d - 1)
        Tx[i, i + 1] = 1
        Tx[i + 1, i] = -1
        Dx = (1 / (2 * dx)) * Tx
        Dx[0, -1] = -1 * (1 / (2 * dx))
        Dx[-1, 0] = 1 * (1 / (2 * dx))
        self.Dx = Dx
        x = np.linspace(-1, 1, num=d)
        u0 = 0.5 * (np.cos(4.5 * np.pi * x) + 1)
        super().__init__(f"Burgers_{d_x}", [0] * d, [1] * d, u0, **kw)

    def f_orig(self, t, u):
        # systems.py:444-450
        return self.Dxx @ u - u * (self.Dx @ u)


def preset(system, N=None):
    """configs.py:6-181 -- per-system (tspan, N, per-slice Ng, Nf, G, F)."""
    if isinstance(system, FHN_ODE):
        N0 = 40
        Ng = N0 * 4
        cfg = dict(tspan=[0, 40], N=N0, Ng=Ng / N0, Nf=int(160000 / 160 * Ng) / N0, G="RK2", F="RK4")
    elif isinstance(system, Rossler):
        cfg = dict(tspan=[0, 340], N=40, Ng=90000 / 40, Nf=4500000 / 40, G="RK1", F="RK4")
    elif isinstance(system, Hopf):
        Ng = 2 * 1024
        cfg = dict(tspan=[-20, 500], N=N, Ng=Ng / N, Nf=Ng * 85 / N, G="RK1", F="RK8")
    elif isinstance(system, DblPend):
        Ng = 3072 + 32
        cfg = dict(tspan=[0, 80], N=32, Ng=Ng / 32, Nf=Ng * 70 / 32, G="RK1", F="RK8")
    elif isinstance(system, Brusselator):
        cfg = dict(tspan=[0, 100], N=25, Ng=10, Nf=1000, G="RK4", F="RK4")
    elif isinstance(system, Lorenz):
        cfg = dict(tspan=[0, 18], N=50, Ng=6, Nf=450, G="RK4", F="RK4")
    elif isinstance(system, FHN_PDE):
        table = {10: (3, 150, "RK2"), 12: (12, 550, "RK2"), 14: (25, 950, "RK2")}
        mul, T, G = table.get(system.d_x, (25, 1100, "RK4"))
        N0 = 512
        Ng = N0 * mul
        Nf = int(np.ceil(1e4 / Ng) * Ng)
        cfg = dict(tspan=[0, T], N=N0, Ng=Ng / N0, Nf=Nf / N0, G=G, F="RK8")
    else:
        raise Exception("No config for input ODE")
    for key in ("N", "Ng", "Nf"):
        cfg[key] = int(cfg[key])
    return cfg
