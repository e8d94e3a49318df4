"""The binding a maintainer of the reference adds to run nnGParareal on libnngpara.so.

Every class below DERIVES FROM THE REFERENCE'S OWN CLASS, so the reference's type checks (`parareal.py:37-41`) and
its unmodified driver `Parareal._parareal` (`parareal.py:212-471`) accept them:

    ODE subclasses      systems.py:23-77   get_vector_field() returns a device field (system id + parameters +
                                           normalisation) instead of a Python closure
    CudaSolverRK        solver.py:72-113   run_F / run_G / run_*_full / run_*_timed on the GPU, + run_F_batch
    CudaNNGP            models.py:97-270   fit keeps the device dataset up to date, predict = one nngp_predict_host
    CudaPool            parareal.py:16-24  executor protocol: the unmodified line parareal.py:311
                                           `pool.map(solver.run_F_timed, ...)` becomes ONE batched launch
    CudaParareal        parareal.py:26-112 the reference's extension idiom (nnGPara_with_time.py:187-215): `_run` builds
                                           the CUDA model, then calls the inherited, unmodified `_parareal`

The arithmetic is not re-implemented here: the mix-in halves of these classes are the package's thin ctypes callers
(nearest_neighbors_gparareal_b200), which in turn only call the C ABI of include/nngpara.h.

    from integration.cuda_backend import bind
    B = bind()                                            # imports the reference from NNGP_REFERENCE_DIR
    ode = B.FHN_PDE(d_x=16); cfg = B.Config(ode, d_x=16).get()
    p = B.CudaParareal(ode, B.CudaSolverRK(ode.get_vector_field(), **cfg), **cfg)
    out = p.run(model='nngp', pool=B.CudaPool(), parall='mpi', nn=20)
"""
import time
import types

import nearest_neighbors_gparareal_b200 as pkg
from nearest_neighbors_gparareal_b200 import systems as pkg_systems

from .ref_env import import_reference

_SYSTEMS = ("FHN_ODE", "Rossler", "Hopf", "DblPend", "Brusselator", "Lorenz", "ThomasLabyrinth", "FHN_PDE", "Burgers")
_bound = None


def bind(ref_dir=None):
    """returns a namespace of classes derived from the reference's (imported from ref_dir / NNGP_REFERENCE_DIR)"""
    global _bound
    if _bound is not None:
        return _bound
    ref = import_reference(ref_dir)
    B = types.SimpleNamespace(ref=ref, Config=ref.configs.Config, CudaPool=pkg.CudaPool)

    class DeviceODE:
        """mix-in for the reference's ODE subclasses: the field is evaluated by libnngpara.so"""
        system_key = None
        device_params = pkg_systems.ODE.device_params
        device_desc = pkg_systems.ODE.device_desc

        def device_system(self, handle=None):
            if not hasattr(self, "_dev"):
                self._dev = {}
            return pkg_systems.ODE.device_system(self, handle)

        def get_vector_field(self):          # systems.py:32-44
            return pkg_systems.DeviceVectorField(self)

    for name in _SYSTEMS:
        ref_cls = getattr(ref.systems, name)
        pkg_cls = getattr(pkg_systems, name)
        body = {"system_key": pkg_cls.system_key, "__doc__": f"reference systems.{name} with a device vector field"}
        if "device_params" in pkg_cls.__dict__:
            body["device_params"] = pkg_cls.__dict__["device_params"]
        setattr(B, name, type(name, (DeviceODE, ref_cls), body))

    class CudaSolverRK(pkg.CudaSolverRK, ref.solver.SolverRK):
        """solver.py:72-113 on the GPU.  The reference's constructor would build two RK objects around a Python
        closure (solver.py:82-83); the package's takes the device field instead and keeps Ng, Nf, F, G, thresh."""

    class CudaNNGP(pkg.CudaNNGP, ref.models.NNGP_p):
        """models.py:97-270: same constructor kwargs, rng stream, timing keys; kNN + fits + selection + mean are one
        C-ABI call (nngp_predict_host)."""

    class CudaParareal(ref.parareal.Parareal):
        """The reference driver, unmodified, with the CUDA model plugged in through `_run`
        (the override idiom of nnGPara_with_time.py:187-215)."""

        def _get_pool(self, *args, **kwargs):          # parareal.py:58-64: an int / None asks for CPU workers
            pool = kwargs.get('pool', None)
            return CudaPool() if (pool is None or isinstance(pool, int)) else pool

        def _run(self, model='nngp', cstm_mdl_name=None, add_model=False, **kwargs):
            if isinstance(model, ref.models.ModelAbstr):
                mdl = model
            elif model.lower() == 'nngp':
                kw = dict(kwargs)
                mdl = CudaNNGP(n=self.n, N=self.N, worker_pool=kw.pop('pool', None), **kw)
            elif model.lower() == 'parareal':
                mdl = ref.models.BareParareal(N=self.N, **kwargs)
            else:
                return super()._run(model=model, cstm_mdl_name=cstm_mdl_name, add_model=add_model, **kwargs)
            s_time = time.time()
            out = self._parareal(mdl, **kwargs)          # parareal.py:212-471, unmodified
            out['timings']['runtime'] = time.time() - s_time
            if add_model:
                out['mdl'] = mdl.store()
            self.runs[mdl.name if cstm_mdl_name is None else cstm_mdl_name] = out
            return out

    CudaPool = pkg.CudaPool
    B.CudaSolverRK, B.CudaNNGP, B.CudaParareal = CudaSolverRK, CudaNNGP, CudaParareal
    # module-level names: the reference pickles the driver and the model copy (parareal.py:114-139, models.py:64-72),
    # and pickle stores classes by module + qualified name
    for cls in [CudaSolverRK, CudaNNGP, CudaParareal] + [getattr(B, n) for n in _SYSTEMS]:
        cls.__module__, cls.__qualname__ = __name__, cls.__name__
        globals()[cls.__name__] = cls
    _bound = B
    return B


def __getattr__(name):
    """unpickling a dump written through this binding looks the classes up here: bind on first use"""
    if name in ("CudaSolverRK", "CudaNNGP", "CudaParareal") + _SYSTEMS:
        bind()
        return globals()[name]
    raise AttributeError(name)
