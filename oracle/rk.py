"""Oracle: explicit fixed-step Runge-Kutta (TEST INFRASTRUCTURE ONLY).

Follows reference RK.py:30-48 (tableaus) and RK.py:113-137 (`_RK_numpy_`, the
NumPy branch taken when use_jax=False), plus RK.py:146-174 for the constant-dt
variant (`_RK_jax_last`).
"""
import numpy as np

_S21 = np.sqrt(21)


def tableau(method):
    """Butcher tableau (a[S,S], b[S], c[S]); reference RK.py:30-48."""
    if method == "RK1":
        a = [[0.0]]
        b = [1.0]
        c = [0.0]
    elif method == "RK2":
        a = [[0, 0], [0.5, 0]]
        b = [0, 1]
        c = [0, 0.5]
    elif method == "RK4":
        a = [[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 1, 0]]
        b = [1 / 6, 1 / 3, 1 / 3, 1 / 6]
        c = [0, 0.5, 0.5, 1]
    elif method == "RK8":
        s = _S21
        a = np.zeros((11, 11))
        a[1, 0] = 1 / 2
        a[2, :2] = [1 / 4, 1 / 4]
        a[3, :3] = [1 / 7, (-7 - 3 * s) / 98, (21 + 5 * s) / 49]
        a[4, :4] = [(11 + s) / 84, 0, (18 + 4 * s) / 63, (21 - s) / 252]
        a[5, :5] = [(5 + s) / 48, 0, (9 + s) / 36, (-231 + 14 * s) / 360, (63 - 7 * s) / 80]
        a[6, :6] = [(10 - s) / 42, 0, (-432 + 92 * s) / 315, (633 - 145 * s) / 90,
                    (-504 + 115 * s) / 70, (63 - 13 * s) / 35]
        a[7, :7] = [1 / 14, 0, 0, 0, (14 - 3 * s) / 126, (13 - 3 * s) / 63, 1 / 9]
        a[8, :8] = [1 / 32, 0, 0, 0, (91 - 21 * s) / 576, 11 / 72, (-385 - 75 * s) / 1152,
                    (63 + 13 * s) / 128]
        a[9, :9] = [1 / 14, 0, 0, 0, 1 / 9, (-733 - 147 * s) / 2205, (515 + 111 * s) / 504,
                    (-51 - 11 * s) / 56, (132 + 28 * s) / 245]
        a[10, :10] = [0, 0, 0, 0, (-42 + 7 * s) / 18, (-18 + 28 * s) / 45, (-273 - 53 * s) / 72,
                      (301 + 53 * s) / 72, (28 - 28 * s) / 45, (49 - 7 * s) / 18]
        b = [1 / 20, 0, 0, 0, 0, 0, 0, 49 / 180, 16 / 45, 49 / 180, 1 / 20]
        c = [0, 1 / 2, 1 / 2, (7 + s) / 14, (7 + s) / 14, 1 / 2, (7 - s) / 14, (7 - s) / 14,
             1 / 2, (7 + s) / 14, 1]
    else:
        raise NotImplementedError("Only RK1, RK2, RK4 and RK8 are implemented")
    return np.array(a, dtype=float), np.array(b, dtype=float), np.array(c, dtype=float)


def rk_last(f, method, t0, t1, steps, u0, h_mode="linspace", _full=False):
    """u(t1) after `steps` explicit RK steps from u(t0)=u0.

    h_mode='linspace' : the NumPy path, RK.py:91-99 + 113-137 -- node times come
        from np.linspace(t0, t1, steps+1) and h = t[n+1]-t[n] varies in its last
        bits from step to step.
    h_mode='const'    : the JAX path, RK.py:101-106 + 146-174 -- h = (t1-t0)/steps,
        t += h.
    Arithmetic order (RK.py:123-135): k_0 = h*f(t,u); for i>=1:
    temp = ((0 + a_i0 k_0) + a_i1 k_1) + ... over ALL j<i (zeros included), k_i =
    h*f(t + c_i h, u + temp); u <- u + np.sum(b*k, axis=1).
    """
    a, b, c = tableau(method)
    steps = int(steps)
    S = b.shape[0]
    u = np.array(u0, dtype=float)
    dim = u.shape[0]
    if h_mode == "linspace":
        tt = np.linspace(t0, t1, num=steps + 1)
    else:
        dt = (t1 - t0) / steps
        t = t0
    brow = b.reshape(1, S)
    traj = [u.copy()] if _full else None
    for n in range(steps):
        if h_mode == "linspace":
            t = tt[n]
            h = tt[n + 1] - tt[n]
        else:
            h = dt
        k = np.zeros((dim, S))
        k[:, 0] = h * f(t, u)
        for i in range(1, S):
            temp = np.zeros(dim)
            for j in range(i):
                temp = temp + a[i, j] * k[:, j]
            k[:, i] = h * f(t + c[i] * h, u + temp)
        u = u + np.sum(brow * k, 1)
        if _full:
            traj.append(u.copy())
        if h_mode != "linspace":
            t = t + dt
    return np.stack(traj) if _full else u


def rk_full(f, method, t0, t1, steps, u0):
    """Every step of the solve, [steps+1, d] -- reference RK.run (RK.py:91-99) with the NumPy branch
    `_RK_numpy_` (RK.py:113-137); what SolverRK.run_F_full / run_G_full return (solver.py:109-113)."""
    return rk_last(f, method, t0, t1, steps, u0, "linspace", _full=True)
