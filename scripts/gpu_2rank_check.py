"""Scratch (torchrun, 2+ ranks): sharded fine solves == unsharded, on every rank; device driver results equal across ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = _lib.default_handle(local)
N = 64
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get(); cfg["N"] = N; cfg["tspan"] = [0, cfg["tspan"][1] * N / 512]; cfg["Nf"] = 200
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
t = np.linspace(cfg["tspan"][0], cfg["tspan"][1], N + 1)
rng = np.random.default_rng(0)
u = ode.get_init_cond()[None, :] + 0.01 * rng.standard_normal((N, 512))
ref = solver.run_F_batch(t[:-1], t[1:], u)
res = nn.CudaPool(sharded=True).map(solver.run_F_timed, t[:-1], t[1:], [u[i] for i in range(N)])
got = np.array([r[0] for r in res])
print(f"rank {rank}: sharded pool == unsharded: {np.array_equal(got, ref)} maxdiff {np.abs(got-ref).max():.3e} finite {np.isfinite(got).all()}", flush=True)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, epsilon=5e-7, verbose="")
out = par.run(model="nngp", nn=12, seed=45, early_stop=2)
v = torch.from_numpy(out["u"]).cuda()
lst = [torch.empty_like(v) for _ in range(world)]
dist.all_gather(lst, v)
print(f"rank {rank}: device driver k={out['k']} conv {out['conv_int']} err {np.nanmax(out['err'],axis=0)} equal across ranks "
      f"{all(torch.equal(lst[0], x) for x in lst)}", flush=True)
dist.destroy_process_group()
