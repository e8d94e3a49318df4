"""Scratch GPU diagnostics: end-to-end runs on the golden cases, printing parity numbers."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn
from helpers import load_run, case_system, device_system

names = sys.argv[1:] or ["lorenz_N32_m11", "lorenz_N50_m11", "lorenz_N50_adaptive", "hopf_N32_m15",
                         "burgers_d32_N32_m12", "fhn_d32_N32_m12"]
for name in names:
    z, cfg, mkw = load_run(name)
    key, kw = case_system(name)
    ode = device_system(key, **kw)
    solver = nn.CudaSolverRK(ode.get_vector_field(), **{k: cfg[k] for k in ("Ng", "Nf", "F", "G")})
    p = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=cfg["N"], epsilon=float(z["epsilon"]), verbose='')
    t = time.time()
    out = p.run(model='nngp', **mkw)
    el = time.time() - t
    # serial fine solution
    N = cfg["N"]
    tt = out['t']
    fine = np.zeros_like(out['u'])
    fine[0] = ode.get_init_cond()
    for i in range(N):
        fine[i + 1] = solver.run_F(tt[i], tt[i + 1], fine[i])
    print(f"{name}: K {out['k']} (ref {int(z['K'])}) conv {out['conv_int']} ref {list(z['conv_int'])} time {el:.2f}s (ref {float(z['seconds']):.0f}s)")
    print(f"    |u-ref_last| {np.max(np.abs(out['u'] - z['u_last'])):.2e}  |u-fine| {np.max(np.abs(out['u']-fine)):.2e}  |ref-fine| {np.max(np.abs(z['u_last']-fine)):.2e}")
    k = min(out['k'], int(z['K']))
    print("    err max/iter gpu", np.array2string(np.nanmax(out['err'], 0)[:k], precision=3))
    print("    err max/iter ref", np.array2string(np.nanmax(z['err'], 0)[:k], precision=3))
    print("    timings", {k_: (round(v, 3) if isinstance(v, float) else None) for k_, v in out['timings'].items() if not hasattr(v, 'shape')})
