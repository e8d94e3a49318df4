"""One GPU emulating the W ranks of the dimension-sharded sweep (nngp_sweep_shard block by block, as
tests/test_gpu_e2e.py::test_dimension_sharded_sweep... does): per-rank GP-fit time of an iteration = total / W.
usage: shard_emulation_probe.py [W=8] [n_slices=120]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 120
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # slices swept (untimed) before the timed ones
N, m = 512, 20
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get(); cfg["Nf"] = 25
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, verbose="")
model = nn.CudaNNGP(n=512, N=N, nn=m, seed=45, handle=h)
st = par.device_setup(model)
par.device_fine_step(st)
I = st["I"]
h.append_iteration(st["u_cur"], st["uF"], st["uG_cur"], N, I, 512, st["stream"])
if skip:
    st0 = torch.from_numpy(model.draw_starts(skip)).cuda()
    h.sweep(st["sys"], st["mG"], solver.h_mode, solver.Ng, st["t"], I + skip, I, m, 1, st0, 0.1, 0.1, st["u_next"], st["uG_next"], 512, st["stream"])
    torch.cuda.synchronize()
    I = I + skip
starts = torch.from_numpy(model.draw_starts(ns)).cuda()
for budget in [int(b) for b in os.environ.get("BUDGETS", "0,100,60").split(",")]:
    h.set_fit_budget(budget)
    u, g = st["u_next"].clone(), st["uG_next"].clone()
    h.profile_read(reset=True); h.profile_enable(True)
    torch.cuda.synchronize(); t0 = time.time()
    for i in range(I, I + ns):
        for rank in range(W):
            j0, dl = nn.parareal.dim_block(512, rank, W)
            h.sweep_shard(st["sys"], st["mG"], solver.h_mode, solver.Ng, st["t"], I + ns, I, i, 1, m, 1, starts, 0.1, 0.1, u, g, 512, j0, dl, st["stream"])
    torch.cuda.synchronize(); wall = time.time() - t0
    h.profile_enable(False)
    prof = h.profile_read(reset=True)
    print(f"W={W} budget={budget}: gp_fit per rank per predict {prof['gp_fit'][0] / W / ns * 1e3:.1f} us "
          f"(x511 = {prof['gp_fit'][0] / W / ns * 511:.1f} ms per iteration), checksum {float(u.sum()):.17g}", flush=True)
