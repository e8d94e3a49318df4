"""Synthetic nnGP kernel sweep (BASELINE.json configs[4]) on one GPU:
kNN scan bandwidth against the measured HBM peak and batched fit throughput.
    python scripts/kernel_sweep.py > profiles/r01/kernel_sweep.log
Data: X ~ U(-1,1)^{n x d}, Y = 1e-3 sin(X W), W ~ N(0,1)/sqrt(d), queries = rows + 1e-3 N(0,1); default_rng(0);
Nelder-Mead starts from default_rng(45)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nearest_neighbors_gparareal_b200 import _lib

h = _lib.default_handle(0)
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5, cold=True):
    best = 1e30
    for _ in range(reps):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print(f"# kNN: argsort(cdist(q, X))[:m], bit-exact strict left-to-right sums; HBM peak {HBM} GB/s (MEASURED_PEAKS.json)")
print("# n      d    Q   m   ms(cold L2)  GB/s(8nd bytes)  frac_HBM   ms(warm)")
rng = np.random.default_rng(0)
for n, d in ((1024, 512), (4096, 512), (16384, 512), (65536, 512), (65536, 128), (65536, 32), (65536, 3), (262144, 128)):
    x = rng.uniform(-1, 1, (n, d)); y = np.zeros_like(x)
    h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, y)
    for Q, m in ((1, 20), (8, 20), (512, 20)):
        if Q * n * 8 > 2 << 30:
            continue
        q = torch.from_numpy(x[:Q] + 1e-3 * rng.standard_normal((Q, d))).to(dev)
        idx = torch.empty((Q, m), dtype=torch.int64, device=dev); dist = torch.empty((Q, m), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: h.knn(q, Q, m, 0, idx, dist, st)
        fn(); torch.cuda.synchronize()
        ms_c, ms_w = timed(fn, cold=True), timed(fn, cold=False)
        by = 8.0 * n * d + 8.0 * Q * d + 16.0 * Q * m
        print(f"{n:7d} {d:4d} {Q:4d} {m:3d}   {ms_c:9.3f}   {by/ms_c/1e6:10.1f}   {by/ms_c/1e6/HBM:8.3f}   {ms_w:8.3f}", flush=True)

print("# fits: 9 Nelder-Mead searches + selection + posterior mean per (query, dim); Q = 512 queries batched")
print("# n      d    m    ms      fits/s      NM-runs/s    nll-evals/s   evals/run")
for n, d, m in ((4096, 3, 10), (4096, 32, 10), (16384, 32, 20), (16384, 128, 20), (16384, 128, 30), (4096, 512, 20), (4096, 32, 5)):
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (n, d)); y = 1e-3 * np.sin(x @ (rng.standard_normal((d, d)) / np.sqrt(d)))
    h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, y)
    Q = 512 if d <= 128 else 64
    q = torch.from_numpy(x[rng.permutation(n)[:Q]] + 1e-3 * rng.standard_normal((Q, d))).to(dev)
    idx = torch.empty((Q, m), dtype=torch.int64, device=dev); dist = torch.empty((Q, m), dtype=torch.float64, device=dev)
    starts = torch.from_numpy(np.random.default_rng(45).integers(-8, 0, (Q, d, 9, 1, 2)).astype(np.int8)).to(dev)
    pred = torch.empty((Q, d), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    h.knn(q, Q, m, 0, idx, dist, st)
    fn = lambda: h.fit_predict(q, idx, dist, Q, m, 1, starts, 0.1, 0.1, pred, stream=st)
    fn(); torch.cuda.synchronize()
    h.counters(reset=True)
    ms = timed(fn, reps=3, cold=False)
    runs, evals = h.counters(reset=True)
    runs /= 3; evals /= 3
    print(f"{n:7d} {d:4d} {m:3d} {ms:9.3f} {Q*d/ms*1e3:11.0f} {runs/ms*1e3:12.0f} {evals/ms*1e3:13.0f} {evals/max(runs,1):8.1f}  finite {bool(torch.isfinite(pred).all())}", flush=True)
