import sys
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from oracle import nngp as onn
from helpers import load_run, samples
name = sys.argv[1]
z, cfg, mkw = load_run(name)
x, D = z["x"], z["D"]
rng = np.random.default_rng(0)
tot = same_traj = same_sel = 0
pred_diff = []
for s in samples(z)[:6]:
    n, m = int(s["n_rows"]), int(s["m"])
    idx, kq = onn.knn(s["query"], x[:n], m)
    r2 = onn.pairwise_sqdist(x[idx], x[idx])
    d = D.shape[1]
    for j in range(min(d, 3)):
        y = D[idx, j]
        res0, res1 = [], []
        for a, jit in enumerate(onn.JITTERS):
            st = s["starts"][j, a, 0].astype(float)
            th0, f0, n0 = onn.nm_run(r2, y, st, jit, 0.1, 0.1)
            f_noisy = lambda th: onn.neg_log_lik(r2, y, th, jit) * (1 + 2.2e-16 * rng.integers(-2, 3))
            from oracle.nelder_mead import nelder_mead
            th1, f1, n1, _, _ = nelder_mead(f_noisy, st, xatol=0.1, fatol=0.1)
            tot += 1
            same_traj += bool(np.array_equal(th0, th1) and n0 == n1)
            res0.append((f0, th0)); res1.append((f1, th1))
        b0 = onn.select([r[0] for r in res0]); b1 = onn.select([r[0] for r in res1])
        p0 = onn.posterior_mean(r2, kq, y, res0[b0][1], onn.JITTERS[b0]); p1 = onn.posterior_mean(r2, kq, y, res1[b1][1], onn.JITTERS[b1])
        same_sel += (b0 == b1)
        pred_diff.append(abs(p0 - p1))
print(name, "runs with identical trajectory under +-2ulp objective noise:", same_traj, "/", tot, " same selection", same_sel, "/", len(pred_diff), " max |dpred|", max(pred_diff), "median", np.median(pred_diff))
