"""BASELINE.json configs[2]: viscous Burgers d=128, T=5.9, N=128 -- Parareal vs GParareal vs nnGParareal on one B200.
Published (Burges_scal_final/*_5.9_128_*, 141 CPU workers): K = 90 / 8 / 14, runtime 33 694 / 12 382 / 5 877 s.
usage: three_way_burgers.py [fine steps per slice, default 40000]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn

Nf = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
rows = []
for model, kw in (("parareal", {}), ("nngp", dict(nn=18, seed=45)), ("gpjax", {})):
    ode = nn.Burgers(d_x=128, normalization='-11')
    solver = nn.CudaSolverRK(ode.get_vector_field(), Ng=4, Nf=Nf, G='RK1', F='RK8')
    drv = nn.PararealDevice if model != "gpjax" else nn.Parareal
    p = drv(ode, solver, tspan=[0, 5.9], N=128, epsilon=5e-7, verbose='')
    t0 = time.time()
    out = p.run(model=model, pool=nn.CudaPool(), parall='mpi', **kw)
    secs = time.time() - t0
    row = dict(model=model, K=out['k'], converged=bool(out['converged']), conv_int=out['conv_int'], runtime_s=round(secs, 2),
               model_s=round(float(out['timings'].get('mdl_tot_t', 0.0)), 2), F_s=round(float(out['timings']['F_time']), 2),
               published_K={"parareal": 90, "nngp": 14, "gpjax": 8}[model],
               published_runtime_s={"parareal": 33694, "nngp": 5877, "gpjax": 12382}[model], fine_steps=Nf)
    print(json.dumps(row), flush=True)
