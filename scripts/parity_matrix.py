"""GPU: convergence parity of every golden run (tests/golden/run_*.npz, recorded from the unmodified reference)
and of the full-size FHN target under different failed-pivot thresholds of the GP factorisation
(nngp_set_pivot_guard).  Prints one JSON line per (case, guard).  usage: parity_matrix.py [guards] [--full] [--dump DIR]"""
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn
from nearest_neighbors_gparareal_b200 import _lib
from helpers import load_run, case_system, device_system

args = [a for a in sys.argv[1:] if not a.startswith("--")]
guards = [float(g) for g in args[0].split(",")] if args else [1.0, 4.0]
full = "--full" in sys.argv
dump_dir = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
h = _lib.default_handle(0)

names = sorted(os.path.basename(f)[4:-4] for f in glob.glob(os.path.join(ROOT, "tests", "golden", "run_*.npz"))
               if "replay" not in f and not f.endswith("_gp.npz"))
for guard in guards:
    h.set_pivot_guard(guard)
    for name in names:
        z, cfg, mkw = load_run(name)
        key, kw = case_system(name)
        ode = device_system(key, **kw)
        solver = nn.CudaSolverRK(ode.get_vector_field(), **{k: cfg[k] for k in ("Ng", "Nf", "F", "G")})
        p = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=cfg["N"], epsilon=float(z["epsilon"]), verbose='')
        t0 = time.time()
        try:
            out = p.run(model='nngp', **mkw)
        except Exception as e:  # noqa
            print(json.dumps(dict(case=name, guard=guard, error=repr(e))), flush=True)
            continue
        eo, er = np.nanmax(out['err'], 0), np.nanmax(z['err'], 0)
        print(json.dumps(dict(case=name, guard=guard, K=out['k'], K_ref=int(z['K']), conv=out['conv_int'],
                              conv_ref=[int(v) for v in z['conv_int']], same=out['conv_int'] == [int(v) for v in z['conv_int']],
                              err=[float('%.4g' % v) for v in eo], err_ref=[float('%.4g' % v) for v in er],
                              du=float(np.max(np.abs(out['u'] - z['u_last']))), secs=round(time.time() - t0, 2))), flush=True)

if full:
    for guard in guards:
        h.set_pivot_guard(guard)
        for norm in (None, "-11"):
            ode = nn.FHN_PDE(d_x=16, normalization=norm)
            cfg = nn.Config(ode, d_x=16).get()
            cfg["Nf"] = 195325
            solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
            par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=cfg["N"], epsilon=5e-7, verbose="")
            hist = dict(u=[], uG=[], uF=[], I=[])

            def hook(k, st, hist=hist):
                if k == 0:
                    hist['u'].append(st['u_cur'].cpu().numpy())
                    hist['uG'].append(st['uG_cur'].cpu().numpy())
                hist['uF'].append(st['uF'][:st['u_cur'].shape[0]].cpu().numpy())
                hist['u'].append(st['u_next'].cpu().numpy())
                hist['uG'].append(st['uG_next'].cpu().numpy())
                hist['I'].append(st['I'])

            dump = dump_dir is not None and guard == guards[0] and norm is None
            t0 = time.time()
            out = par.run(model="nngp", nn=20, seed=45, iteration_hook=hook if dump else None)
            print(json.dumps(dict(case="fhn_d512_N512_m20_full", guard=guard, normalization=norm, K=out['k'],
                                  conv=out['conv_int'], err=[float(v) for v in np.nanmax(out['err'], 0)],
                                  argmax=[int(np.nanargmax(out['err'][:, c])) for c in range(out['err'].shape[1])],
                                  secs=round(time.time() - t0, 2))), flush=True)
            if dump:
                os.makedirs(dump_dir, exist_ok=True)
                kmax = min(4, len(hist['uF']))
                np.savez(os.path.join(dump_dir, f"fhn_full_state_{'id' if norm is None else 'n11'}.npz"),
                         u=np.stack(hist['u'][:kmax + 1]), uG=np.stack(hist['uG'][:kmax + 1]),
                         uF=np.stack(hist['uF'][:kmax]), I_sweep=np.array(hist['I'][:kmax]),
                         conv_int=np.array(out['conv_int']), err=out['err'], t=out['t'], guard=guard)
