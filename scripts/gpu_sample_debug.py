"""Scratch: per-run comparison GPU vs oracle on predict calls recorded from the reference run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from nearest_neighbors_gparareal_b200 import _lib
from oracle import nngp as onn
from helpers import load_run, samples
np.set_printoptions(linewidth=200, precision=6)
h = _lib.default_handle(0)
name = sys.argv[1] if len(sys.argv) > 1 else "hopf_N32_m15"
maxs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
z, cfg, mkw = load_run(name)
x, D = z["x"], z["D"]
d = x.shape[1]
h.dataset_reset(); h.dataset_reserve(x.shape[0], d); h.dataset_append_host(x, D)
for s in samples(z)[:maxs]:
    n, m = int(s["n_rows"]), int(s["m"])
    R = s["starts"].shape[2]
    out = h.predict_host(s["query"][None], m, s["starts"][None], R, 0.1, 0.1, n_rows=n, details=True)
    opred, od = onn.predict(s["query"], x[:n], D[:n], m, s["starts"].astype(np.int64), return_details=True)
    print(f"call {int(s['call'])} k={int(s['k'])} i={int(s['i'])} n={n} m={m}  |gpu-ref| {np.abs(out['pred'][0]-s['preds'])}  |oracle-ref| {np.abs(opred-s['preds'])} preds {s['preds']}")
    for j in range(min(d, 3)):
        print(f"  dim {j}: gpu sel theta {out['theta_opt'][0,j]} jit {out['jitter_opt'][0,j]} f {out['fval_opt'][0,j]:.10f} | oracle theta {od['theta_opt'][j]} jit {od['jitter_opt'][j]} f {od['fval_opt'][j]:.10f}")
        for a in range(9):
            g_th, g_f, g_n = out['thetas'][0, j, a, 0], out['fvals'][0, j, a, 0], out['nfev'][0, j, a, 0]
            o_th, o_f, o_n = od['thetas'][j, a, 0], od['fvals'][j, a, 0], od['nfev'][j, a, 0]
            flag = "==" if (np.array_equal(g_th, o_th) and g_n == o_n) else "!="
            pg = onn.posterior_mean(od['r2'], od['dist'], D[od['idx'], j], g_th, -20.0 + a)
            po = onn.posterior_mean(od['r2'], od['dist'], D[od['idx'], j], o_th, -20.0 + a)
            print(f"     jit {-20+a} start {s['starts'][j,a,0]} {flag} gpu {g_th} f {g_f:.12f} n {g_n} pred {pg:.6e} | ora {o_th} f {o_f:.12f} n {o_n} pred {po:.6e}")
