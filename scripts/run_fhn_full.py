"""Full nnGParareal solve of the FHN-PDE target (d=512, N=512, m=20) on the device; compares K, conv_int and
the per-iteration error maxima with the reference's published run (FHN_scal_times_16_512_nngp)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nearest_neighbors_gparareal_b200 as nn

fine = int(sys.argv[1]) if len(sys.argv) > 1 else 195325
norm = sys.argv[2] if len(sys.argv) > 2 else None
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 45
ode = nn.FHN_PDE(d_x=16, normalization=norm)
cfg = nn.Config(ode, d_x=16).get()
cfg["Nf"] = fine
solver = nn.CudaSolverRK(ode.get_vector_field(), **cfg)
par = nn.PararealDevice(ode, solver, tspan=cfg["tspan"], N=cfg["N"], epsilon=5e-7, verbose="")
t = time.time()
out = par.run(model="nngp", nn=20, seed=seed)
el = time.time() - t
pub = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "published.json")))
ref = pub["FHN_scal_times/FHN_scal_times_16_512_nngp"]["NNGP"]
res = {"fine_steps": fine, "normalization": norm, "seed": seed, "K": out["k"], "conv_int": out["conv_int"],
       "converged": bool(out["converged"]), "err_max_per_iter": [float(v) for v in np.nanmax(out["err"], 0)],
       "runtime_s": el, "timings": {k: float(v) for k, v in out["timings"].items() if isinstance(v, float)},
       "published": {"K": ref["K"], "conv_int": ref["conv_int"], "err_max_per_iter": ref["err_max_per_iter"],
                     "runtime_s": 17849.0}}
print(json.dumps(res))
