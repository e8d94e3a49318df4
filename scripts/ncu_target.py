"""Small profiling target: FHN-PDE d=512, N=512 state after the coarse initialisation, then the first
few slices of the sweep (each = G launch, kNN, neighbour matrix, GP fit+predict).  Used under ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib

n_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fine = int(sys.argv[2]) if len(sys.argv) > 2 else 25
N, m = 512, 20
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get()
cfg["Nf"] = fine
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, verbose="")
model = nn.CudaNNGP(n=512, N=N, nn=m, seed=45, handle=h)
st = par.device_setup(model)
par.device_fine_step(st)
h.append_iteration(st["u_cur"], st["uF"], st["uG_cur"], N, st["I"], 512, st["stream"])
starts = torch.from_numpy(model.draw_starts(n_slices)).cuda()
# sweep only the first n_slices slices: N_eff = I + n_slices
h.sweep(st["sys"], st["mG"], solver.h_mode, solver.Ng, st["t"], st["I"] + n_slices, st["I"], m, 1, starts,
        0.1, 0.1, st["u_next"], st["uG_next"], 512, st["stream"])
torch.cuda.synchronize()
print("ok", h.counters())
if os.environ.get("NFEV_HIST"):
    import numpy as np
    q = st["u_next"][st["I"] + n_slices - 1].cpu().numpy()
    out = h.predict_host(q[None], m, model.draw_starts(1), 1, 0.1, 0.1, details=True)
    nf = out["nfev"].ravel()
    print("nfev: mean", nf.mean(), "median", np.median(nf), "p90", np.percentile(nf, 90), "p99", np.percentile(nf, 99),
          "max", nf.max(), "count>=200", int((nf >= 200).sum()), "count==400", int((nf >= 400).sum()))
    fv = out["fvals"].ravel()
    print("nfev==400 with finite fval", int(((nf >= 400) & np.isfinite(fv)).sum()), "nfev>=150 finite", int(((nf >= 150) & np.isfinite(fv)).sum()),
          "nfev>=100 finite", int(((nf >= 100) & np.isfinite(fv)).sum()), "max nfev among finite", int(nf[np.isfinite(fv)].max()))
    print("nfev histogram (finite searches):", np.histogram(nf[np.isfinite(fv)], bins=[0, 30, 50, 70, 100, 150, 200, 300, 401])[0])
    print("inf fvals", int(np.isinf(fv).sum()), "of", fv.size)
    long = np.argsort(nf)[-8:]
    for t in long:
        print("  task", t, "dim", t // 9, "jit", -20 + t % 9, "nfev", nf[t], "f", fv[t], "theta", out["thetas"].reshape(-1, 2)[t])
