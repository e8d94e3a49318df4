"""Scratch: FHN d=512 N=512 iteration 0 on the device; find slices with large corrections and inspect them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
np.set_printoptions(linewidth=200, precision=5)
N, m = 512, 20
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get()
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, verbose="")
model = nn.CudaNNGP(n=512, N=N, nn=m, seed=45, handle=h)
st = par.device_setup(model)
par.device_fine_step(st)
par.device_sweep(st, 0)
err = par.device_errors(st)
print("err max", np.nanmax(err), "argmax", int(np.nanargmax(err)), "first >0.05:", np.flatnonzero(err > 0.05)[:10])
print("err[1:12]", err[1:12])
D = (st["uF"][1:N + 1] - st["uG_cur"][1:N + 1]).cpu().numpy()
print("|F-G| rows max", np.abs(D).max(), "typical", np.median(np.abs(D)))
i = int(np.flatnonzero(err > 0.05)[0]) - 1 if np.any(err > 0.05) else 5   # slice whose predict produced u_next[i+1]
q = st["u_next"][i].cpu().numpy()
pred_used = (st["u_next"][i + 1] - st["uG_next"][i + 1]).cpu().numpy()
print("slice", i, "max |pred|", np.abs(pred_used).max(), "at dim", int(np.abs(pred_used).argmax()), " |F-G| of that slice", np.abs(D[i]).max())
out = h.predict_host(q[None], m, model.draw_starts(1), 1, 0.1, 0.1, details=True)
j = int(np.abs(pred_used).argmax())
print("idx", out["idx"][0])
print("dim", j, "pred(now, other starts)", out["pred"][0, j], "theta", out["theta_opt"][0, j], "jit", out["jitter_opt"][0, j], "f", out["fval_opt"][0, j])
print("fvals", out["fvals"][0, j].ravel())
print("thetas", out["thetas"][0, j].reshape(-1, 2))
print("nfev", out["nfev"][0, j].ravel())
print("y", D[out["idx"][0], j])

# ---- the same predict with the SAME starts as the sweep used, on the device and on the oracle
from oracle import nngp as onn
model2 = nn.CudaNNGP(n=512, N=N, nn=m, seed=45, handle=h)
all_starts = model2.draw_starts(N - 1)
s_i = all_starts[i - 1]            # predict number (i - I) with I = 1
out = h.predict_host(q[None], m, s_i[None], 1, 0.1, 0.1, details=True)
print("\nsame starts: device pred", out["pred"][0, j], "theta", out["theta_opt"][0, j], "jit", out["jitter_opt"][0, j], "f", out["fval_opt"][0, j])
print("device fvals", out["fvals"][0, j].ravel())
print("device thetas", out["thetas"][0, j].reshape(-1, 2))
X = st["u_cur"][0:N].cpu().numpy()
oi, okq = onn.knn(q, X, m)
print("oracle idx equal", np.array_equal(oi, out["idx"][0]))
r2 = onn.pairwise_sqdist(X[oi], X[oi])
yj = D[oi, j]
res = [onn.nm_run(r2, yj, s_i[j, a, 0].astype(float), onn.JITTERS[a], 0.1, 0.1) for a in range(9)]
print("oracle fvals", np.array([r[1] for r in res]))
print("oracle thetas", np.array([r[0] for r in res]))
b = onn.select([r[1] for r in res])
print("oracle pred", onn.posterior_mean(r2, okq, yj, res[b][0], onn.JITTERS[b]), "selected", b)
print("oracle posterior mean at the device's theta", onn.posterior_mean(r2, okq, yj, out["theta_opt"][0, j], out["jitter_opt"][0, j]))
K = onn.se_kernel_from_r2(r2, out["theta_opt"][0, j]) + np.eye(m) * 10 ** out["jitter_opt"][0, j]
print("cond(K) at device theta", np.linalg.cond(K), "r2 range", r2[r2 > 0].min(), r2.max())
