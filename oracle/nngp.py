"""Oracle: nearest-neighbour GP correction model (TEST INFRASTRUCTURE ONLY).

NumPy/SciPy restatement of reference models.py `NNGP_p` (:97-270) and
`_fit_gp_jit` (:86-92).  Third-party arithmetic the reference relies on:
scipy.spatial.distance.cdist('sqeuclidean') (a strict left-to-right, non-fused
sum of (q_j-x_j)^2), numpy argsort, LAPACK potrf via
linalg.cholesky, SciPy solve_triangular, SciPy Nelder-Mead (oracle/nelder_mead.py).
"""
import itertools

import numpy as np
import scipy.linalg

from .nelder_mead import nelder_mead

JITTERS = np.arange(-20, -11, dtype=float)  # models.py:186


def sqdist_rows(q, x):
    """||q - x_i||^2 for every row, summed strictly left to right without FMA --
    the arithmetic of SciPy's cdist 'sqeuclidean' used at models.py:177."""
    q = np.asarray(q, dtype=float).ravel()
    x = np.asarray(x, dtype=float)
    s = np.zeros(x.shape[0])
    for j in range(x.shape[1]):
        diff = q[j] - x[:, j]
        s = s + diff * diff
    return s


def knn(q, x, m):
    """models.py:177-179: indices of the m nearest rows in ascending distance.
    Ties are broken by index (stable sort) -- the contract of BASELINE.json."""
    dist = sqdist_rows(q, x)
    order = np.argsort(dist, kind="stable")[:m]
    return order.astype(np.int64), dist[order]


def pairwise_sqdist(a, b):
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    return np.stack([sqdist_rows(ai, b) for ai in a])


def se_kernel_from_r2(r2, theta):
    """models.py:145-148: k = 10^sy * exp(-0.5 * (1/10^sx) * r2)."""
    sigma_x, sigma_y = theta
    with np.errstate(all="ignore"):
        return 10 ** (sigma_y) * np.exp(-0.5 * (1 / (10 ** sigma_x)) * r2)


def fit_gp(r2, y, theta, jitter):
    """models.py:86-92: K = k(x,x) + I*10^jitter, L = chol(K) (NaN on failure, as
    XLA's Cholesky does), alpha = L^-T (L^-1 y)."""
    m = r2.shape[0]
    with np.errstate(all="ignore"):
        K = se_kernel_from_r2(r2, theta) + np.eye(m) * 10 ** jitter
    try:
        if not np.all(np.isfinite(K)):
            raise np.linalg.LinAlgError
        L = np.linalg.cholesky(K)
    except np.linalg.LinAlgError:
        return None, None
    z = scipy.linalg.solve_triangular(L, y, lower=True, check_finite=False)
    alpha = scipy.linalg.solve_triangular(L.T, z, lower=False, check_finite=False)
    return L, alpha


def neg_log_lik(r2, y, theta, jitter):
    """models.py:240-252: 0.5 y.alpha + sum(log diag L) + (m/2) log 2pi; NaN -> +inf."""
    L, alpha = fit_gp(r2, y, theta, jitter)
    if L is None:
        return np.inf
    m = y.shape[0]
    with np.errstate(all="ignore"):
        res = -(-0.5 * y.T @ alpha - np.sum(np.log(np.diag(L))) - (m / 2) * np.log(2 * np.pi))
    if np.isnan(res):
        return np.inf
    return float(res)


def draw_starts(rng, d, n_restarts):
    """models.py:190-192: one rng.integers(-8, 0, 2) per (dim, jitter, restart) task, in
    itertools.product order."""
    n_tasks = d * JITTERS.shape[0] * n_restarts
    return np.stack([rng.integers(-8, 0, 2) for _ in range(n_tasks)]).reshape(
        d, JITTERS.shape[0], n_restarts, 2)


def nm_run(r2, y, start, jitter, fatol, xatol):
    """models.py:228-237, 254-260: one Nelder-Mead search from an integer start."""
    x, fval, nfev, _, _ = nelder_mead(lambda th: neg_log_lik(r2, y, th, jitter), start,
                                      xatol=xatol, fatol=fatol)
    return x, fval, nfev


def select(fvals):
    """models.py:212-215: mask fval < 0.9*min; if empty use all; first minimum in
    (jitter, restart) order."""
    fvals = np.asarray(fvals, dtype=float).ravel()
    with np.errstate(all="ignore"):
        mask = fvals < fvals.min() * 0.9
    if mask.sum() == 0:
        mask[:] = True
    idx = np.flatnonzero(mask)
    return int(idx[int(np.argmin(fvals[idx]))])


def posterior_mean(r2, kq, y, theta, jitter):
    """models.py:162-168: refit at theta, mean = k(xm, q)^T alpha."""
    L, alpha = fit_gp(r2, y, theta, jitter)
    if L is None:
        return np.nan
    kstar = se_kernel_from_r2(kq, theta)
    return float(kstar.T @ alpha)


def predict(q, x, y, m, starts, fatol=1e-1, xatol=1e-1, return_details=False):
    """models.py:171-226 for one query.  `starts` is [d, 9, R, 2] (draw_starts)."""
    q = np.asarray(q, dtype=float).ravel()
    idx, kq = knn(q, x, m)
    xm, ym = x[idx], y[idx]
    r2 = pairwise_sqdist(xm, xm)
    d = y.shape[1]
    R = starts.shape[2]
    preds = np.empty(d)
    theta_opt = np.empty((d, 2))
    jit_opt = np.empty(d)
    fval_opt = np.empty(d)
    fvals = np.empty((d, JITTERS.shape[0], R))
    thetas = np.empty((d, JITTERS.shape[0], R, 2))
    nfev = np.zeros((d, JITTERS.shape[0], R), dtype=np.int32)
    for j in range(d):
        for a, jit in enumerate(JITTERS):
            for r in range(R):
                th, fv, ne = nm_run(r2, ym[:, j], starts[j, a, r], jit, fatol, xatol)
                thetas[j, a, r], fvals[j, a, r], nfev[j, a, r] = th, fv, ne
        best = select(fvals[j])
        a, r = divmod(best, R)
        theta_opt[j], jit_opt[j], fval_opt[j] = thetas[j, a, r], JITTERS[a], fvals[j, a, r]
        preds[j] = posterior_mean(r2, kq, ym[:, j], theta_opt[j], jit_opt[j])
    if return_details:
        return preds, dict(idx=idx, dist=kq, r2=r2, theta_opt=theta_opt, jitter_opt=jit_opt,
                           fval_opt=fval_opt, fvals=fvals, thetas=thetas, nfev=nfev)
    return preds


class OracleNNGP:
    """Model protocol of models.py:19-72, 97-226 (fit / predict / get_times)."""

    name = "NNGP"

    def __init__(self, n, N, nn="adaptive", n_restarts=1, seed=45, fatol=None, xatol=None, **_):
        self.n, self.N = n, N
        self.nn, self.n_restarts = nn, n_restarts
        self.fatol = 1e-1 if fatol is None else fatol
        self.xatol = 1e-1 if xatol is None else xatol
        self.rng = np.random.default_rng(seed)
        self.nm_runs = 0
        self.nm_evals = 0

    def fit(self, x, y, k):
        self.x, self.y, self.k = x, y, k

    def predict(self, new_x, prev_F=None, prev_G=None, i=None, return_details=False):
        m = max(10, self.k + 2) if self.nn == "adaptive" else self.nn
        starts = draw_starts(self.rng, self.n, self.n_restarts)
        out = predict(new_x, self.x, self.y, m, starts, self.fatol, self.xatol, return_details=True)
        self.nm_runs += out[1]["nfev"].size
        self.nm_evals += int(out[1]["nfev"].sum())
        return out if return_details else out[0]


class OracleBareParareal:
    """models.py:74-83."""
    name = "Parareal"

    def __init__(self, **_):
        pass

    def fit(self, *a, **k):
        pass

    def predict(self, new_x, prev_F, prev_G, **_):
        return prev_F - prev_G
