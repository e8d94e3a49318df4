"""GPU probe: round / packing statistics of the grouped search kernel and first-failing-pivot histogram.
Needs a statistics build of gpfit.cu (-DNNGP_FIT_STATS, see profiles/r02/fit_failing_pivots.log) loaded through NNGPARA_LIB;
with the normal build it only prints the search and evaluation counts."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
N, m = 512, 20
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get()
cfg["Nf"] = 25
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, verbose="")
model = nn.CudaNNGP(n=512, N=N, nn=m, seed=45, handle=h)
st = par.device_setup(model)
par.device_fine_step(st)
h.append_iteration(st["u_cur"], st["uF"], st["uG_cur"], N, st["I"], 512, st["stream"])
ns = 80
starts = torch.from_numpy(model.draw_starts(ns)).cuda()
I = st["I"]
for lo, hi in ((0, 2), (2, 50), (50, 80)):
    h.counters(reset=True)
    h.sweep(st["sys"], st["mG"], solver.h_mode, solver.Ng, st["t"], I + hi, I + lo, m, 1, starts[lo:hi].contiguous(), 0.1, 0.1, st["u_next"], st["uG_next"], 512, st["stream"])
    torch.cuda.synchronize()
    print("slices", lo, hi, flush=True)
    h.counters(reset=True)
    try:
        h.lib.nngp_fit_stats_dump()
    except AttributeError:
        pass
