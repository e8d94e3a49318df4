"""Oracle: the nnGParareal outer loop (TEST INFRASTRUCTURE ONLY).

Restates reference parareal.py:212-471 (`Parareal._parareal`) in the O(N d)
memory form of `PararealLight._parareal` (:812-1060), which the reference states
is equivalent.  Plain serial NumPy; no timing tables, plots or checkpoints.
"""
import numpy as np

from .rk import rk_last


class OracleSolver:
    """solver.py:72-107 (`SolverRK`): Ng / Nf steps per slice with methods G / F."""

    def __init__(self, f, Ng, Nf, F, G, h_mode="linspace"):
        self.f, self.Ng, self.Nf, self.F, self.G, self.h_mode = f, int(Ng), int(Nf), F, G, h_mode

    def run_F(self, t0, t1, u0):
        return rk_last(self.f, self.F, t0, t1, self.Nf, u0, self.h_mode)

    def run_G(self, t0, t1, u0):
        return rk_last(self.f, self.G, t0, t1, self.Ng, u0, self.h_mode)


def parareal(u0, solver, tspan, N, model, epsilon=5e-7, early_stop=None, hook=None):
    """Returns dict(t, u, err, x, D, k, converged, conv_int); parareal.py:469-471."""
    u0 = np.asarray(u0, dtype=float)
    n = u0.shape[0]
    t = np.linspace(tspan[0], tspan[1], num=N + 1)
    I = 0
    conv_int = []
    err = np.full((N + 1, N), np.nan)
    u_cur = np.full((N + 1, n), np.nan)
    uG_cur = np.full((N + 1, n), np.nan)
    uF_cur = np.full((N + 1, n), np.nan)
    u_cur[0] = uG_cur[0] = uF_cur[0] = u0
    # parareal.py:264-277 -- serial coarse initialisation
    for i in range(N):
        uG_cur[i + 1] = solver.run_G(t[i], t[i + 1], uG_cur[i])
    u_cur[:] = uG_cur
    u_next, uG_next, uF_next = u_cur.copy(), uG_cur.copy(), uF_cur.copy()
    u_next[1:] = np.nan
    uG_next[1:] = np.nan
    x = np.zeros((0, n))
    D = np.zeros((0, n))
    k = 0
    for k in range(N):
        # parareal.py:309-329 -- fine solves of the unconverged slices
        for i in range(I, N):
            uF_cur[i + 1] = solver.run_F(t[i], t[i + 1], u_cur[i])
        # :331-334 -- slice I+1 is now exact
        uG_next[I + 1] = uG_cur[I + 1]
        uF_next[I + 1] = uF_cur[I + 1]
        u_next[I + 1] = uF_cur[I + 1]
        I += 1
        # :336-339 -- dataset append
        x = np.vstack([x, u_cur[I - 1:N]])
        D = np.vstack([D, uF_cur[I:N + 1] - uG_cur[I:N + 1]])
        if I == N:
            # :343-348
            err[:, k] = np.linalg.norm(u_next - u_cur, np.inf, 1)
            err[-1, k] = np.nextafter(epsilon, 0)
            break
        model.fit(x, D, k=k)
        # :359-382 -- serial sweep
        for i in range(I, N):
            uG_next[i + 1] = solver.run_G(t[i], t[i + 1], u_next[i])
            preds = model.predict(u_next[i].reshape(1, -1), uF_cur[i + 1], uG_cur[i + 1], i=i)
            u_next[i + 1] = preds + uG_next[i + 1]
        if np.any(np.isnan(uG_next)):
            raise Exception("NaN values in initial coarse solve - increase Ng!")
        # :402-416 -- convergence bookkeeping
        err[:, k] = np.linalg.norm(u_next - u_cur, np.inf, 1)
        err[I, k] = 0
        u_cur[:] = u_next
        uG_cur[:] = uG_next
        for p in range(I + 1, N + 1):
            if err[p, k] < epsilon:
                uF_next[p] = uF_cur[p]
                I += 1
            else:
                break
        uF_cur[:] = uF_next
        conv_int.append(I)
        if hook is not None:
            hook(k=k, I=I, u=u_cur, x=x, D=D, err=err[:, k])
        if I == N:
            break
        if early_stop is not None and k == early_stop - 1:
            break
    return dict(t=t, u=u_cur.copy(), err=err[:, :k + 1], x=x, D=D, k=k + 1,
                converged=(I == N), conv_int=conv_int)
