"""Scratch: fine-step time against the number of slices on one GPU for the two FHN kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
h = _lib.default_handle(0)
ode = nn.FHN_PDE(d_x=16)
cfg = nn.Config(ode, d_x=16).get(); cfg["Nf"] = 20000
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
hh, sysid = solver.device()
rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
for n in (32, 64, 128, 148, 256, 296, 384, 444, 512, 592):
    u0 = torch.from_numpy(ode.get_init_cond()[None, :] + 0.01 * rng.standard_normal((n, 512))).to(dev)
    u1 = torch.empty_like(u0)
    t0 = torch.zeros(n, dtype=torch.float64, device=dev); t1 = t0 + 2.0
    row = [f"n={n:4d}"]
    for tile in ("0", "1", "2"):
        os.environ["NNGP_RK_TILE"] = tile
        st = torch.cuda.current_stream().cuda_stream
        hh.rk_batch(sysid, 8, solver.h_mode, 2000, n, t0, t1, u0, 512, u1, 512, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); hh.rk_batch(sysid, 8, solver.h_mode, 20000, n, t0, t1, u0, 512, u1, 512, st); e1.record()
        torch.cuda.synchronize()
        row.append(f"{ {'0': 'point', '1': '2x2', '2': '1x2'}[tile]} {e0.elapsed_time(e1)*195325/20000:7.1f} ms")
    print("  ".join(row), flush=True)
