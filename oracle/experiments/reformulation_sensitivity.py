"""Does evaluating the data-fit term as z.z (z = L^-1 y) instead of y.alpha change the run?"""
import sys, time, json
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, scipy.linalg
from oracle import nngp as onn, parareal as opara
from helpers import load_run, oracle_system, case_system

def nll_ztz(r2, y, theta, jitter):
    m = r2.shape[0]
    with np.errstate(all="ignore"):
        K = onn.se_kernel_from_r2(r2, theta) + np.eye(m) * 10 ** jitter
    try:
        if not np.all(np.isfinite(K)): raise np.linalg.LinAlgError
        L = np.linalg.cholesky(K)
    except np.linalg.LinAlgError:
        return np.inf
    z = scipy.linalg.solve_triangular(L, y, lower=True, check_finite=False)
    with np.errstate(all="ignore"):
        res = -(-0.5 * (z @ z) - np.sum(np.log(np.diag(L))) - (m / 2) * np.log(2 * np.pi))
    return np.inf if np.isnan(res) else float(res)

name = sys.argv[1]
variant = sys.argv[2]
if variant == "ztz":
    onn.neg_log_lik = nll_ztz
z, cfg, mkw = load_run(name)
key, kw = case_system(name)
s = oracle_system(key, **kw)
solver = opara.OracleSolver(s.f, cfg["Ng"], cfg["Nf"], cfg["F"], cfg["G"])
model = onn.OracleNNGP(n=s.dim(), N=cfg["N"], **mkw)
t = time.time()
out = opara.parareal(s.u0, solver, cfg["tspan"], cfg["N"], model, epsilon=float(z["epsilon"]))
print(name, variant, "K", out["k"], "conv", out["conv_int"], "ref K", int(z["K"]), list(map(int, z["conv_int"])), f"{time.time()-t:.0f}s",
      "errmax", np.array2string(np.nanmax(out["err"], 0), precision=3), flush=True)
