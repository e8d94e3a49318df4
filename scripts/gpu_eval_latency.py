"""Scratch: latency of one objective evaluation in a lone warp and throughput with every SM sub-partition loaded
(gp_nll kernel: one warp per (query, dim), nt evaluations each)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nearest_neighbors_gparareal_b200 import _lib
h = _lib.default_handle(0)
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d = 4
n = 4096
x = rng.uniform(-1, 1, (n, d)); y = 1e-3 * np.sin(x @ rng.standard_normal((d, d)))
h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, y)
for nq, nt in ((1, 4000), (37, 4000), (148, 2000), (592, 2000), (1184, 1000), (2368, 1000)):
    Q = x[:nq] + 1e-3
    idx, dist = h.knn_host(Q, m)
    theta = np.stack([rng.uniform(-3, 0, (nq, d, nt)), rng.uniform(-6, -1, (nq, d, nt))], axis=-1)
    j10 = np.full((nq, d, nt), 1e-16)
    t_idx = torch.from_numpy(idx).to(dev); t_th = torch.from_numpy(theta).to(dev); t_j = torch.from_numpy(j10).to(dev)
    out = torch.empty((nq, d, nt), dtype=torch.float64, device=dev)
    h.gp_nll(t_idx, nq, m, nt, t_th, t_j, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.gp_nll(t_idx, nq, m, nt, t_th, t_j, out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    warps = nq * d
    print(f"m={m} warps={warps:5d} ({warps/592:.2f}/SMSP) nt={nt}: {ms:.3f} ms -> {ms*1e-3*1.965e9/nt:.0f} cycles per eval per warp, "
          f"{warps*nt/ms/1e3:.2f} M evals/s, finite {float(torch.isfinite(out).double().mean()):.3f}")
