"""TEST INFRASTRUCTURE ONLY -- experiment, not a test.

Which failure rule of the device factorisation agrees with LAPACK's potrf (what the
reference reaches through models.py:86-92) on the nearly singular neighbour matrices
of a steady state?  Builds an FHN-PDE dataset that reaches its steady state (rows
become bitwise identical), takes the m nearest neighbours of late-slice queries and
evaluates the objective on a grid of (theta, jitter) with

  lapack   np.linalg.cholesky (the oracle / the reference under the shim)
  guard4   right-looking LDL^T with FMA, pivot <= 4 ulp * K_rr fails (round 1 rule)
  plain    same factorisation, pivot <= 0 fails
  snap     same factorisation, pivots and finalised column entries rounded to the
           ulp(K_rr) grid (what a sum-first dot product does), pivot <= 0 fails

and prints the agreement of the +inf sets and of the finite values.

usage: python -m oracle.experiments.pivot_rule_study [golden-run.npz]
"""
import sys
from fractions import Fraction

import numpy as np

from .. import nngp as onn
from .. import rk as ork
from .. import systems as osys


def fma(a, b, c):
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return a * b + c
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def ldl_device(K, y, rule):
    """lane-per-row right-looking LDL^T as csrc/gpfit.cu::gp_core; returns nll without the constant or inf"""
    M = K.shape[0]
    a = K.copy()
    dd = np.array([K[r, r] for r in range(M)])
    dd0 = dd[0]
    z = y.astype(float).copy()
    pmin = dd0 * 8.8817841970012523e-16 if rule == "guard4" else 0.0
    quad, logdet = 0.0, 0.0

    def snap(v):
        return (v + dd0) - dd0

    for k in range(M):
        p = dd[k]
        if rule == "snap":
            p = snap(p)
        if not (p > pmin):
            return np.inf
        ip = 1.0 / p
        quad = fma(z[k] * z[k], ip, quad)
        logdet += np.log(p)
        for r in range(k + 1, M):
            ark = a[r, k]
            if rule == "snap":
                ark = snap(ark)
            w = ark * ip
            dd[r] = fma(-w, ark, dd[r])
            z[r] = fma(-w, z[k], z[r])
            for j in range(k + 1, r):
                ajk = snap(a[j, k]) if rule == "snap" else a[j, k]
                a[r, j] = fma(-w, ajk, a[r, j])
    return 0.5 * (quad + logdet)


def nll_lapack(K, y):
    try:
        L = np.linalg.cholesky(K)
    except np.linalg.LinAlgError:
        return np.inf
    import scipy.linalg
    zz = scipy.linalg.solve_triangular(L, y, lower=True, check_finite=False)
    v = 0.5 * zz @ zz + np.sum(np.log(np.diag(L)))
    return v if v == v else np.inf


def steady_dataset(d_x=4, N=96, T=206.25):
    sy = osys.FHN_PDE(d_x=d_x)
    t = np.linspace(0, T, N + 1)
    G = lambda a, b, u: ork.rk_last(sy.f, "RK4", a, b, 25, u)
    F = lambda a, b, u: ork.rk_last(sy.f, "RK8", a, b, 25, u)
    u = [sy.u0]
    for i in range(N):
        u.append(G(t[i], t[i + 1], u[-1]))
    u = np.array(u)
    xs, ys = [], []
    for it in range(3):  # plain parareal iterations build the (u, F-G) pairs
        uF = np.array([F(t[i], t[i + 1], u[i]) for i in range(N)])
        uG = np.array([G(t[i], t[i + 1], u[i]) for i in range(N)])
        xs.append(u[:-1].copy())
        ys.append(uF - uG)
        un = u.copy()
        for i in range(N):
            un[i + 1] = G(t[i], t[i + 1], un[i]) + uF[i] - uG[i]
        u = un
    return np.vstack(xs), np.vstack(ys), u


def main():
    rng = np.random.default_rng(0)
    if len(sys.argv) > 1:
        z = np.load(sys.argv[1])
        x, y = z["x"], z["D"]
        qs = [z["u_last"][i] for i in (z["u_last"].shape[0] * np.array([0.3, 0.5, 0.7, 0.9])).astype(int)]
        m = 20
    else:
        x, y, u = steady_dataset()
        qs = [u[i] for i in (30, 50, 70, 90)]
        m = 20
    rules = ("guard4", "plain", "snap")
    tot = {r: dict(both_inf=0, both_fin=0, dev_inf_only=0, ref_inf_only=0, maxrel=0.0) for r in rules}
    for q in qs:
        idx, kq = onn.knn(q, x, m)
        r2 = onn.pairwise_sqdist(x[idx], x[idx])
        ndup = int(np.sum(r2[np.triu_indices(m, 1)] == 0.0))
        print("query: zero-distance pairs among neighbours:", ndup, "min r2>0:",
              np.min(r2[r2 > 0]) if np.any(r2 > 0) else None, flush=True)
        for j in rng.choice(y.shape[1], 3, replace=False):
            yy = y[idx, j]
            for _ in range(40):
                th = rng.uniform(-8.5, 0.5, 2)
                jit = float(rng.integers(-20, -11))
                with np.errstate(all="ignore"):
                    K = onn.se_kernel_from_r2(r2, th) + np.eye(m) * 10 ** jit
                ref = nll_lapack(K, yy)
                for r in rules:
                    dev = ldl_device(K, yy, r)
                    s = tot[r]
                    if np.isinf(ref) and np.isinf(dev):
                        s["both_inf"] += 1
                    elif np.isinf(dev):
                        s["dev_inf_only"] += 1
                    elif np.isinf(ref):
                        s["ref_inf_only"] += 1
                    else:
                        s["both_fin"] += 1
                        s["maxrel"] = max(s["maxrel"], abs(dev - ref) / max(1.0, abs(ref)))
    for r in rules:
        print(r, tot[r])


if __name__ == "__main__":
    main()
