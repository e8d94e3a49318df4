"""Small target for compute-sanitizer (racecheck / memcheck): one launch of every shared-memory kernel of the path --
the 2x2-tile FHN propagator (barrier per RK stage), the point-per-thread PDE propagator, the fused sweep prologue
(last-CTA-done ticket), both Nelder-Mead search kernels (per-warp shared tiles, __syncwarp) and selection / mean."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib

h = _lib.default_handle(0)
rng = np.random.default_rng(0)
ode = nn.FHN_PDE(d_x=16)
s = nn.CudaSolverRK(ode.get_vector_field(), Ng=2, Nf=3, F='RK8', G='RK4')
u0 = ode.get_init_cond()[None] + 0.01 * rng.standard_normal((3, ode.get_dim()))
print("fhn tile", np.isfinite(s.run_F_batch([0, 1, 2], [1, 2, 3], u0)).all(), flush=True)
b = nn.Burgers(d_x=32, normalization='-11')
sb = nn.CudaSolverRK(b.get_vector_field(), Ng=2, Nf=3, F='RK8', G='RK1')
print("burgers", np.isfinite(sb.run_F_batch([0, 0.1], [0.1, 0.2], np.stack([b.get_init_cond()] * 2))).all(), flush=True)
n, d, m = 96, 4, 20
x = rng.uniform(-1, 1, (n, d)); x[5:8] = x[4]
y = 1e-3 * np.sin(x @ rng.standard_normal((d, d)))
h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, y)
Q = x[[4, 20, 33, 50, 70]] + 1e-3 * rng.standard_normal((5, d))
st = rng.integers(-8, 0, (5, d, 9, 1, 2)).astype(np.int8)
for mode in ("warp", "grouped"):
    h.set_fit_mode(mode)
    out = h.predict_host(Q, m, st, 1, 0.1, 0.1)
    print("fit", mode, np.isfinite(out["pred"]).all(), flush=True)
h.set_fit_mode("auto")
# fused sweep prologue through the device sweep of a tiny run
o = nn.FHN_PDE(d_x=4)
cfg = nn.Config(o, d_x=4).get(); cfg["N"] = 6; cfg["tspan"] = [0, 12]
sv = nn.CudaSolverRK(o.get_vector_field(), **cfg)
out = nn.PararealDevice(o, sv, tspan=cfg["tspan"], N=6, verbose='').run(model='nngp', nn=5, early_stop=2)
print("sweep", out["k"], flush=True)
