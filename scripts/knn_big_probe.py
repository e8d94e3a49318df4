import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nearest_neighbors_gparareal_b200 import _lib
h = _lib.default_handle(0); dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
n, d, Q, m = 65536, 512, 1, 20
x = rng.uniform(-1, 1, (n, d)); h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, np.zeros_like(x))
q = torch.from_numpy(x[:Q] + 1e-3).to(dev)
idx = torch.empty((Q, m), dtype=torch.int64, device=dev); dist = torch.empty((Q, m), dtype=torch.float64, device=dev)
for _ in range(3):
    h.knn(q, Q, m, 0, idx, dist, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("ok")
