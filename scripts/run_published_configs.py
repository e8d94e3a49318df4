"""Full-size runs of the other BASELINE.json configurations on the device, against the reference's published K
(tests/golden/published.json): Burgers d=128 N=128 (configs[2]) and Hopf N=32..512 at the preset fine-step counts
(configs[1]; the published runs used Nf x 1e4 + paging, so K may differ by +-1)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn

pub = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "published.json")))
rows = []
ode = nn.Burgers(d_x=128, normalization='-11')
cfg = nn.Config(ode).get()
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=cfg["N"], epsilon=5e-7, verbose="")
t = time.time(); out = par.run(model="nngp", nn=18, seed=45); el = time.time() - t
ref = pub["Burges_scal_final/Burges_scal_final_5.9_128_nngp"]["NNGP"]
rows.append(dict(config="Burgers d=128 N=128 m=18 F=RK8x40000", K=out["k"], conv_int=out["conv_int"], runtime_s=round(el, 2),
                 published_K=ref["K"], published_conv_int=ref["conv_int"], published_runtime_s=5877))
for N in (32, 64, 128, 256, 512):
    ode = nn.Hopf(normalization='-11')
    cfg = nn.Config(ode, N=N).get()
    solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
    par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=N, epsilon=5e-7, verbose="")
    t = time.time(); out = par.run(model="nngp", nn=15, n_restarts=2, seed=45); el = time.time() - t
    ref = pub[f"nonaut_scal_final/nonaut_scal_final_{N}_nngp"]["NNGP"]
    rows.append(dict(config=f"Hopf N={N} m=15 R=2 (preset Nf)", K=out["k"], conv_int=out["conv_int"], runtime_s=round(el, 2),
                     published_K=ref["K"], published_conv_int=ref["conv_int"]))
for r in rows:
    print(json.dumps(r))
