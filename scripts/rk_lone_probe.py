"""Scratch: fine step with one slice per SM or fewer (the multi-GPU rank case), for ncu source-level stalls."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nearest_neighbors_gparareal_b200 as nn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get(); cfg["Nf"] = steps
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
h, sysid = solver.device()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
u0 = torch.from_numpy(ode.get_init_cond()[None, :] + 0.01 * rng.standard_normal((n, 512))).to(dev)
u1 = torch.empty_like(u0)
t0 = torch.zeros(n, dtype=torch.float64, device=dev); t1 = t0 + 2.0
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    h.rk_batch(sysid, 8, solver.h_mode, steps, n, t0, t1, u0, 512, u1, 512, st)
torch.cuda.synchronize(); print("ok")
