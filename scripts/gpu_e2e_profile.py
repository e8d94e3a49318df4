"""Scratch: cProfile of the host-protocol sweep (run_G_timed + predict_timed per slice) at the FHN target."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
N, m = 512, 20
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get(); cfg["Nf"] = 2000
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
t = np.linspace(cfg["tspan"][0], cfg["tspan"][1], N + 1)
d = 512
u_cur = np.empty((N + 1, d)); u_cur[0] = ode.get_init_cond()
for i in range(N):
    u_cur[i + 1] = solver.run_G(t[i], t[i + 1], u_cur[i])
uG_cur = u_cur.copy()
uF = np.empty((N + 1, d)); uF[0] = u_cur[0]
uF[1:] = solver.run_F_batch(t[:-1], t[1:], u_cur[:-1])
def sweep(n_slices):
    model = nn.CudaNNGP(n=d, N=N, nn=m, seed=45, handle=h)
    u_next, uG_next = u_cur.copy(), uG_cur.copy()
    u_next[1] = uF[1]
    model.fit_timed(u_cur[0:N], uF[1:N + 1] - uG_cur[1:N + 1], k=0)
    tg = tp = 0.0
    for i in range(1, 1 + n_slices):
        a = time.perf_counter()
        uG_next[i + 1], _ = solver.run_G_timed(t[i], t[i + 1], u_next[i])
        b = time.perf_counter()
        preds = model.predict_timed(u_next[i].reshape(1, -1), uF[i + 1], uG_cur[i + 1], i=i)
        c = time.perf_counter()
        u_next[i + 1] = preds + uG_next[i + 1]
        tg += b - a; tp += c - b
    return tg / n_slices, tp / n_slices
sweep(20)
h.profile_read(reset=True); h.profile_enable(True)
tg, tp = sweep(200)
h.profile_enable(False)
pr = h.profile_read(reset=True)
print(f"per slice: run_G_timed {tg*1e6:.0f} us, predict_timed {tp*1e6:.0f} us; device time per slice: " +
      ", ".join(f"{k} {v[0]/200*1e3:.0f} us" for k, v in pr.items()))
cProfile.run("sweep(100)", "/tmp/e2e.prof")
pstats.Stats("/tmp/e2e.prof").sort_stats("tottime").print_stats(14)
