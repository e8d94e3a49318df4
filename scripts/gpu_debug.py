"""Scratch GPU diagnostics (not a test): compares each kernel with the oracle and prints numbers."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
from oracle import nngp as onn, rk as ork, systems as osys

h = _lib.default_handle(0)
print("fp64 peak TF", h.bench_fp64(20000), "copy GB/s", h.bench_copy(1 << 30))
G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "rk_vectors.npz"))
mk = {"lorenz": lambda: nn.Lorenz(normalization='-11'), "lorenz_id": lambda: nn.Lorenz(),
      "hopf": lambda: nn.Hopf(normalization='-11'), "rossler": lambda: nn.Rossler(normalization='-11'),
      "fhn_ode": lambda: nn.FHN_ODE(normalization='-11'), "brusselator": lambda: nn.Brusselator(normalization='-11'),
      "dblpend": lambda: nn.DblPend(normalization='-11'), "thomas": lambda: nn.ThomasLabyrinth(normalization='-11'),
      "burgers128": lambda: nn.Burgers(d_x=128, normalization='-11'), "burgers32": lambda: nn.Burgers(d_x=32, normalization='-11'),
      "fhn16": lambda: nn.FHN_PDE(d_x=16), "fhn4": lambda: nn.FHN_PDE(d_x=4), "fhn4_n": lambda: nn.FHN_PDE(d_x=4, normalization='-11')}
for name, ctor in mk.items():
    ode = ctor()
    U = G[f"{name}_u"]
    f = ode.get_vector_field()
    got = f(0.3, U)
    want = G[f"{name}_f"]
    line = f"{name:12s} f: maxrel {np.max(np.abs(got-want)/(np.abs(want)+1e-300)):.2e} exact {np.array_equal(got, want)}"
    for method in ("RK1", "RK2", "RK4", "RK8"):
        t0, t1, steps = G[f"{name}_{method}_t"]
        s = nn.CudaSolverRK(f, Ng=int(steps), Nf=int(steps), F=method, G=method)
        got = s.run_F_batch([t0] * 3, [t1] * 3, U[:3])
        want = G[f"{name}_{method}_u1"]
        line += f" | {method} {np.max(np.abs(got-want)/(np.abs(want)+1e-300)):.1e} {'=' if np.array_equal(got, want) else '~'}"
    print(line)

# kNN + fit
rng = np.random.default_rng(1)
for (n, d, m) in ((300, 3, 11), (2000, 32, 12), (3000, 512, 20)):
    x = rng.uniform(-1, 1, (n, d))
    y = 1e-3 * np.sin(x @ (rng.standard_normal((d, d)) / np.sqrt(d)))
    Q = x[rng.permutation(n)[:16]] + 1e-3 * rng.standard_normal((16, d))
    h.dataset_reset(); h.dataset_reserve(n, d); h.dataset_append_host(x, y)
    idx, dist = h.knn_host(Q, m)
    ok = True
    for qi in range(16):
        oi, od = onn.knn(Q[qi], x, m)
        ok &= np.array_equal(oi, idx[qi]) and np.array_equal(od, dist[qi])
    print(f"knn n={n} d={d} m={m}: bit-exact {ok}")
    # one predict
    dd = min(d, 8)
    model = nn.CudaNNGP(n=d, N=4, nn=m, seed=45)
    model.fit(x, y, k=0)
    state = model.rng.bit_generator.state
    t = time.time()
    pred, det = model.predict(Q[0].reshape(1, -1), None, None, i=0, return_details=True)
    el = time.time() - t
    orng = np.random.default_rng(); orng.bit_generator.state = state
    starts = onn.draw_starts(orng, d, 1)
    # oracle on the first dd dims only (python is slow)
    oidx, okq = onn.knn(Q[0], x, m)
    r2 = onn.pairwise_sqdist(x[oidx], x[oidx])
    nmatch = ntot = 0
    worst = 0.0
    for j in range(dd):
        for a, jit in enumerate(onn.JITTERS):
            th, fv, ne = onn.nm_run(r2, y[oidx, j], starts[j, a, 0], jit, 0.1, 0.1)
            g_th, g_fv, g_ne = det['thetas'][0, j, a, 0], det['fvals'][0, j, a, 0], det['nfev'][0, j, a, 0]
            same = np.array_equal(th, g_th) and ne == g_ne
            nmatch += same; ntot += 1
            if not same:
                print("   mismatch j", j, "jit", jit, "start", starts[j, a, 0], "oracle", th, fv, ne, "gpu", g_th, g_fv, g_ne)
            elif np.isfinite(fv):
                worst = max(worst, abs(fv - g_fv) / max(abs(fv), 1e-300))
    print(f"predict n={n} d={d} m={m}: {el*1e3:.1f} ms, NM runs identical {nmatch}/{ntot}, worst fval rel {worst:.2e}, mean nfev {det['nfev'].mean():.1f} max {det['nfev'].max()}")
    for j in range(dd):
        om = onn.posterior_mean(r2, okq, y[oidx, j], det['theta_opt'][0, j], det['jitter_opt'][0, j])
        print(f"   dim {j}: pred gpu {pred[j]:.12e} oracle@same-theta {om:.12e} rel {abs(pred[j]-om)/abs(om):.2e} theta {det['theta_opt'][0,j]} jit {det['jitter_opt'][0,j]}")
