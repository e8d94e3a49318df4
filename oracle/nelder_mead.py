"""Oracle: Nelder-Mead simplex search (TEST INFRASTRUCTURE ONLY).

The reference calls SciPy (pinned scipy==1.12.0, requirements.txt:6) at
models.py:257-259: minimize(f, x0, method='Nelder-Mead', options={'fatol','xatol'}).
SciPy is not part of /root/reference, so this restates the published algorithm
of scipy.optimize._optimize._minimize_neldermead (non-adaptive: rho=1, chi=2,
psi=sigma=1/2; initial simplex x0 and x0 with one coordinate times 1.05 (or
0.00025 when it is 0); maxiter = maxfev = 200*N; convergence test at the top of
every iteration; stable re-sort of the simplex after every iteration; an
evaluation requested once maxfev calls were made aborts the iteration without
applying its update).  tests/test_oracle_golden.py checks it bit-for-bit against
the installed scipy.optimize.minimize on the reference's own objective.
"""
import numpy as np


class _Budget(Exception):
    pass


def nelder_mead(func, x0, xatol=1e-4, fatol=1e-4):
    """Returns (x_best, f_best, nfev, nit, status); status 0 ok, 1 maxfev, 2 maxiter."""
    x0 = np.asarray(x0, dtype=float).ravel()
    n = x0.shape[0]
    maxfun = maxiter = 200 * n
    ncalls = [0]

    def f(x):
        if ncalls[0] >= maxfun:
            raise _Budget()
        ncalls[0] += 1
        return float(func(np.copy(x)))

    sim = np.empty((n + 1, n))
    sim[0] = x0
    for k in range(n):
        y = x0.copy()
        y[k] = (1 + 0.05) * y[k] if y[k] != 0 else 0.00025
        sim[k + 1] = y
    fsim = np.full(n + 1, np.inf)
    try:
        for k in range(n + 1):
            fsim[k] = f(sim[k])
    except _Budget:
        pass
    order = np.argsort(fsim, kind="stable")
    sim, fsim = sim[order], fsim[order]

    it = 1
    while ncalls[0] < maxfun and it < maxiter:
        try:
            if (np.max(np.abs(sim[1:] - sim[0])) <= xatol
                    and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol):
                break
            xbar = np.add.reduce(sim[:-1], 0) / n
            worst = sim[-1]
            xr = 2 * xbar - 1 * worst
            fxr = f(xr)
            shrink = False
            if fxr < fsim[0]:
                xe = 3 * xbar - 2 * worst
                fxe = f(xe)
                if fxe < fxr:
                    sim[-1], fsim[-1] = xe, fxe
                else:
                    sim[-1], fsim[-1] = xr, fxr
            elif fxr < fsim[-2]:
                sim[-1], fsim[-1] = xr, fxr
            elif fxr < fsim[-1]:
                xc = 1.5 * xbar - 0.5 * worst
                fxc = f(xc)
                if fxc <= fxr:
                    sim[-1], fsim[-1] = xc, fxc
                else:
                    shrink = True
            else:
                xcc = 0.5 * xbar + 0.5 * worst
                fxcc = f(xcc)
                if fxcc < fsim[-1]:
                    sim[-1], fsim[-1] = xcc, fxcc
                else:
                    shrink = True
            if shrink:
                for j in range(1, n + 1):
                    sim[j] = sim[0] + 0.5 * (sim[j] - sim[0])
                    fsim[j] = f(sim[j])
            it += 1
        except _Budget:
            pass
        order = np.argsort(fsim, kind="stable")
        sim, fsim = sim[order], fsim[order]

    status = 1 if ncalls[0] >= maxfun else (2 if it >= maxiter else 0)
    return sim[0].copy(), float(np.min(fsim)), ncalls[0], it, status
