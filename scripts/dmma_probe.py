"""north_star: "the ||x||^2 - 2 x.y + ||y||^2 term on FP64 DMMA tensor cores only at PDE-scale dimension d, where ncu shows
it pays".  Measurement that decides it: the exact kNN of this package (strict left-to-right (q_j - x_j)^2 sums, bit-equal to
SciPy's cdist) against the expansion form evaluated with an FP64 GEMM (cuBLAS DGEMM = DMMA on B200, through torch.mm) plus
a top-(m + slack) pre-filter, at the shapes of BASELINE configs 4 and 5.  The GEMM is a PROBE of what a DMMA pre-filter
could save, not a product path: the expansion form is not bit-exact (and mis-orders near-duplicate rows), so it can only
pre-filter candidates for an exact re-rank."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nearest_neighbors_gparareal_b200 import _lib

h = _lib.default_handle(0)
dev = torch.device("cuda", 0)


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


rng = np.random.default_rng(0)
print("# n      d    Q   m | exact kNN ms | DGEMM expansion ms (gemm + norms) | + topk(m+12) ms | same top-m set | exact TFLOP/s (3Qnd) | gemm TFLOP/s (2Qnd)")
for n, d, Q, near_dup in ((65536, 512, 512, False), (65536, 128, 512, False), (16384, 512, 512, False), (3055, 512, 1, True),
                          (65536, 512, 1, False), (65536, 512, 16, False)):
    x = rng.uniform(-1, 1, (n, d))
    if near_dup:   # steady-state-like dataset: most rows equal up to 1e-15
        x[n // 10:] = x[n // 10] + 1e-15 * rng.standard_normal((n - n // 10, d))
    m = 20
    h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, np.zeros_like(x))
    qh = x[rng.permutation(n)[:Q]] + 1e-3 * rng.standard_normal((Q, d))
    if near_dup:
        qh = x[n // 2:n // 2 + 1] + 1e-16
    q = torch.from_numpy(qh).to(dev)
    X = torch.from_numpy(x).to(dev)
    idx = torch.empty((Q, m), dtype=torch.int64, device=dev); dist = torch.empty((Q, m), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    f_exact = lambda: h.knn(q, Q, m, 0, idx, dist, st)
    f_exact(); torch.cuda.synchronize()
    t_exact = timed(f_exact)
    xn = (X * X).sum(1)

    def f_gemm():
        g = torch.mm(q, X.T)
        return (q * q).sum(1, keepdim=True) - 2 * g + xn[None]
    f_gemm(); torch.cuda.synchronize()
    t_gemm = timed(f_gemm)
    d2 = f_gemm()
    f_topk = lambda: torch.topk(d2, m + 12, dim=1, largest=False)
    t_topk = timed(f_topk)
    cand = f_topk()[1]
    exact_sets = idx.cpu().numpy()
    cand_np = cand.cpu().numpy()
    contained = np.mean([set(exact_sets[i]) <= set(cand_np[i]) for i in range(Q)])
    top_m_same = np.mean([set(exact_sets[i]) == set(cand_np[i][:m]) for i in range(Q)])
    print(f"{n:7d} {d:4d} {Q:4d} {m:3d} | {t_exact:9.3f} | {t_gemm:9.3f} | {t_topk:8.3f} | exact top-m inside the m+12 candidates for "
          f"{contained:.3f} of the queries, identical top-m for {top_m_same:.3f} | {3.0*Q*n*d/t_exact/1e9:7.2f} | {2.0*Q*n*d/t_gemm/1e9:7.2f}"
          + ("   [near-duplicate rows]" if near_dup else ""), flush=True)
