"""Scratch: cycles per objective evaluation inside the Nelder-Mead fit kernel for a handful of lone searches
(d small -> 9*d searches, each on its own warp), from the kernel time and the longest search."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nearest_neighbors_gparareal_b200 import _lib
h = _lib.default_handle(0)
rng = np.random.default_rng(0)
m = 20
for d in (1, 4, 16, 64, 128, 256, 512):
    n = 600
    x = rng.uniform(-1, 1, (n, d)); y = 1e-9 * np.sin(x @ rng.standard_normal((d, d)))
    h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, y)
    q = x[:1] + 1e-3
    starts = rng.integers(-8, 0, (1, d, 9, 1, 2)).astype(np.int8)
    h.predict_host(q, m, starts, 1, 0.1, 0.1, details=True)
    h.profile_read(reset=True); h.profile_enable(True)
    out = h.predict_host(q, m, starts, 1, 0.1, 0.1, details=True)
    h.profile_enable(False)
    pr = h.profile_read(reset=True)
    nf = out["nfev"].ravel()
    ms = pr["gp_fit"][0]
    print(f"d={d:4d} searches={nf.size:5d} nfev max {nf.max()} mean {nf.mean():.1f} sum {nf.sum()}  fit {ms:.3f} ms -> "
          f"{ms*1e-3*1.965e9/nf.max():.0f} cycles per eval of the longest search; "
          f"{ms*1e-3*1.965e9*592/nf.sum():.0f} SMSP-cycles per eval")
