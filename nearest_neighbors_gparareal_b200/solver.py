"""Host mirror of the reference's solver protocol (solver.py:21-113).

`CudaSolverRK` keeps the constructor of `SolverRK(f, Ng, Nf, F, G, thresh=1e7, use_jax=True)`;
`f` is the callable returned by `ODE.get_vector_field()` (a DeviceVectorField).  Every
propagation is nngp_rk_batch[_host] (csrc/rk.cu); `run_F_batch` advances all slices of an
iteration in ONE launch -- what `pool.map(solver.run_F_timed, ...)` (parareal.py:311) becomes
through `CudaPool`.
"""
import time

import numpy as np

from . import _lib
from .systems import DeviceVectorField, ODE


def calc_time(f):
    def wrapper(*args, **kwargs):
        s_time = time.time()
        ret = f(*args, **kwargs)
        el_time = time.time() - s_time
        return ret, el_time
    return wrapper


class SolverAbstr:
    '''All methods return the ODE solution at time t1 given u0 at time t0 (solver.py:29-69).'''

    def run_F(self, t0, t1, u0):
        raise NotImplementedError('run_F not implemented')

    @calc_time
    def run_F_timed(self, t0, t1, u0):
        return self.run_F(t0, t1, u0)

    def run_F_full(self, t0, t1, u0):
        raise NotImplementedError('run_F_full not implemented')

    @calc_time
    def run_F_full_timed(self, t0, t1, u0):
        return self.run_F_full(t0, t1, u0)

    def run_G(self, t0, t1, u0):
        raise NotImplementedError('run_G not implemented')

    @calc_time
    def run_G_timed(self, t0, t1, u0):
        return self.run_G(t0, t1, u0)

    def run_G_full(self, t0, t1, u0):
        raise NotImplementedError('run_G_full not implemented')

    @calc_time
    def run_G_full_timed(self, t0, t1, u0):
        return self.run_G_full(t0, t1, u0)


class CudaSolverRK(SolverAbstr):
    """solver.py:72-113 on the device.  h_mode: how the step sizes are formed -- 'linspace' = differences of
    np.linspace(t0, t1, steps + 1) (RK.py:101-109 with use_jax=False, the path the golden fixtures were recorded
    on, hence the default) or 'const' = (t1 - t0) / steps for every step (the jitted fori_loop of RK.py:146-203,
    the reference's use_jax=True default).  The two differ in the last bits of h only."""

    def __init__(self, f, Ng, Nf, F, G, thresh=1e7, use_jax=True, h_mode='linspace', handle=None, **kwargs):
        if isinstance(f, ODE):
            f = f.get_vector_field()
        if not isinstance(f, DeviceVectorField):
            raise Exception('f must come from ODE.get_vector_field() of this package: the device '
                            'solver cannot run an arbitrary Python closure (no CPU fallback)')
        for name in (F, G):
            if name not in _lib.METHODS:
                raise NotImplementedError('Only RK1, RK2, RK4 and RK8 are implemented')
        self.f = f
        self.ode = f.ode
        self.Ng = int(Ng)
        self.Nf = int(Nf)
        self.F = F
        self.G = G
        self.thresh = thresh
        self.h_mode = {'linspace': _lib.H_LINSPACE, 'const': _lib.H_CONST}[h_mode]
        self._handle = handle

    # -- device plumbing -----------------------------------------------------------------
    def device(self):
        return self.ode.device_system(self._handle)

    def _batch(self, method, steps, t0, t1, u0):
        h, sys = self.device()
        t0 = np.asarray(t0, dtype=float).ravel()
        t1 = np.asarray(t1, dtype=float).ravel()
        u0 = np.asarray(u0, dtype=float).reshape(t0.shape[0], self.ode.get_dim())
        if t0.shape[0] == 0:
            return u0.copy()  # no slices left (all converged): nothing to launch
        steps = int(steps)
        if steps > self.thresh:
            # solver.py:89-96 (paging quirk kept: every page integrates with the TOTAL step count)
            thresh = int(self.thresh)
            st = steps - 1
            pages = [thresh] * int(st / thresh) + [st % thresh] * (st % thresh != 0)
            step = (t1 - t0) / st
            for page in pages:
                t1p = t0 + step * page
                u0 = h.rk_batch_host(sys, _lib.METHODS[method], self.h_mode, st, t0, t1p, u0)
                t0 = t1p
            return u0
        return h.rk_batch_host(sys, _lib.METHODS[method], self.h_mode, steps, t0, t1, u0)

    def run_F_batch(self, t0, t1, u0):
        """F for many slices at once: u0[n,d] -> u1[n,d]"""
        return self._batch(self.F, self.Nf, t0, t1, u0)

    def run_G_batch(self, t0, t1, u0):
        return self._batch(self.G, self.Ng, t0, t1, u0)

    # -- reference protocol --------------------------------------------------------------
    def run_F(self, t0, t1, u0):
        return self._batch(self.F, self.Nf, [t0], [t1], np.asarray(u0, dtype=float)[None, :])[0]

    def run_G(self, t0, t1, u0):
        return self._batch(self.G, self.Ng, [t0], [t1], np.asarray(u0, dtype=float)[None, :])[0]

    def run_F_full(self, t0, t1, u0):
        """solver.py:109-110: every step of the fine solve, ndarray[Nf+1, d]"""
        h, sys = self.device()
        return h.rk_full_host(sys, _lib.METHODS[self.F], self.h_mode, self.Nf, t0, t1, u0)

    def run_G_full(self, t0, t1, u0):
        """solver.py:112-113"""
        h, sys = self.device()
        return h.rk_full_host(sys, _lib.METHODS[self.G], self.h_mode, self.Ng, t0, t1, u0)

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_handle'] = None
        return state


SolverRK = CudaSolverRK
