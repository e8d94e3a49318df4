"""Scratch: per-search objective-evaluation counts (nfev) of real FHN-PDE d=512 predicts, with the start points,
to study the tail of the persistent fit kernel (gpurun_out/nfev_stats.npz)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
N, m = 512, 20
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get()
cfg["Nf"] = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, verbose="")
model = nn.CudaNNGP(n=512, N=N, nn=m, seed=45, handle=h)
st = par.device_setup(model)
par.device_fine_step(st)
par.device_sweep(st, 0)
torch.cuda.synchronize()
u_next = st["u_next"].cpu().numpy()
rng = np.random.default_rng(7)
rec = {}
for i in (1, 2, 50, 200, 400, 510):
    starts = rng.integers(-8, 0, (1, 512, 9, 1, 2)).astype(np.int8)
    out = h.predict_host(u_next[i][None], m, starts, 1, 0.1, 0.1, details=True)
    nf = out["nfev"].reshape(512, 9)
    rec[f"nfev_{i}"] = nf
    rec[f"starts_{i}"] = starts.reshape(512, 9, 2)
    rec[f"fvals_{i}"] = out["fvals"].reshape(512, 9)
    rec[f"thetas_{i}"] = out["thetas"].reshape(512, 9, 2)
    f = nf.ravel()
    print(f"slice {i}: mean {f.mean():.1f} median {np.median(f)} p90 {np.percentile(f,90)} p99 {np.percentile(f,99)} max {f.max()} "
          f"n>=200 {int((f>=200).sum())} n>=399 {int((f>=399).sum())}")
    # correlation with the start cell
    s = starts.reshape(-1, 2).astype(int)
    tab = np.zeros((8, 8)); cnt = np.zeros((8, 8))
    np.add.at(tab, (s[:, 0] + 8, s[:, 1] + 8), f); np.add.at(cnt, (s[:, 0] + 8, s[:, 1] + 8), 1)
    print(np.round(tab / np.maximum(cnt, 1)).astype(int))
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "nfev_stats.npz"), **rec)
