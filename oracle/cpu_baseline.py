"""CPU baseline sampler -- TEST / BENCH INFRASTRUCTURE ONLY (used by bench.py's cpu_baseline leg
and by `bench.py --impl reference`).

Times the oracle restatement of the reference's CPU path (NumPy/SciPy, the reference's own
process-pool parallelism: one task per Nelder-Mead search, models.py:197, and one task per fine
slice, parareal.py:311) on a bounded sample of the FHN-PDE nnGParareal iteration, on all host
cores, and extrapolates to the full iteration with the reference's own cost model
(article_lib.py:58-115):  T_iter = ceil(N/C) * t_F + (N-1) * (t_G + t_predict(C)).
"""
import math
import multiprocessing as mp
import os
import time

import numpy as np

from . import nngp as onn
from . import rk as ork
from . import systems as osys

_G = {}


def _init(d_x, normalization):
    _G["sys"] = osys.FHN_PDE(d_x=d_x, normalization=normalization)


def _fine(args):
    method, t0, t1, steps, u0 = args
    s = time.perf_counter()
    u1 = ork.rk_last(_G["sys"].f, method, t0, t1, steps, u0)
    return u1, time.perf_counter() - s


def _nm_dim(args):
    """all 9*R searches of one output dimension (models.py:228-237 tasks of that dimension)"""
    r2, y, starts, fatol, xatol = args
    s = time.perf_counter()
    nfev = 0
    for a, jit in enumerate(onn.JITTERS):
        for r in range(starts.shape[1]):
            _, _, ne = onn.nm_run(r2, y, starts[a, r].astype(float), jit, fatol, xatol)
            nfev += ne
    return nfev, time.perf_counter() - s


class FhnCpuSampler:
    """Builds the iteration-0 state of the FHN-PDE run with the oracle (coarse sweep + fine solves
    with `data_fine_steps`), then times samples of it."""

    def __init__(self, d_x=16, N=512, m=20, T=1100.0, g_method="RK4", g_steps=25, f_method="RK8",
                 data_fine_steps=25, cores=None, seed=45, normalization=None):
        self.d_x, self.N, self.m = d_x, N, m
        self.cores = cores or os.cpu_count()
        self.sys = osys.FHN_PDE(d_x=d_x, normalization=normalization)
        self.d = self.sys.dim()
        self.t = np.linspace(0, T, N + 1)
        self.g_method, self.g_steps, self.f_method = g_method, g_steps, f_method
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_init, initargs=(d_x, normalization))
        self.rng = np.random.default_rng(seed)
        # coarse sweep (parareal.py:264-277), serial by nature
        s = time.perf_counter()
        uG = np.empty((N + 1, self.d))
        uG[0] = self.sys.u0
        for i in range(N):
            uG[i + 1] = ork.rk_last(self.sys.f, g_method, self.t[i], self.t[i + 1], g_steps, uG[i])
        self.t_G = (time.perf_counter() - s) / N
        self.uG = uG
        # dataset of iteration 0: x = u^0, D = F(u^0) - G(u^0) (values only; cost is timed separately)
        res = self.pool.map(_fine, [(f_method, self.t[i], self.t[i + 1], data_fine_steps, uG[i]) for i in range(N)])
        uF = np.stack([r[0] for r in res])
        self.x = uG[:N].copy()
        self.D = uF - uG[1:]

    def close(self):
        self.pool.close()
        self.pool.join()

    def sample(self, n_dims, n_slices, fine_steps, sample_fine_steps=25, n_predicts=8):
        """One bounded sample.  Returns a dict with the extrapolated iteration time."""
        N, C = self.N, self.cores
        # fine propagator: n_slices tasks of sample_fine_steps, scaled linearly in the step count
        idx = np.linspace(0, N - 1, n_slices).astype(int)
        s = time.perf_counter()
        res = self.pool.map(_fine, [(self.f_method, self.t[i], self.t[i + 1], sample_fine_steps, self.uG[i]) for i in idx])
        wall_F = time.perf_counter() - s
        t_F_slice = float(np.mean([r[1] for r in res])) * (fine_steps / sample_fine_steps)
        # n_predicts predicts (models.py:171-226; SURVEY.md section 8d asks for >= 8) with queries spread over the slices,
        # each restricted to n_dims / n_predicts output dimensions, the searches farmed over the pool as the reference does
        per = max(1, n_dims // n_predicts)
        tasks, t_knn = [], 0.0
        for i_q in np.linspace(1, N - 1, n_predicts).astype(int):
            q = self.uG[i_q] + 1e-6 * self.rng.standard_normal(self.d)
            s = time.perf_counter()
            nn_idx, kq = onn.knn(q, self.x, self.m)
            r2 = onn.pairwise_sqdist(self.x[nn_idx], self.x[nn_idx])
            t_knn += (time.perf_counter() - s) / n_predicts
            dims = self.rng.permutation(self.d)[:per]
            starts = self.rng.integers(-8, 0, (per, 9, 1, 2))
            tasks += [(r2, self.D[nn_idx, j], starts[k], 0.1, 0.1) for k, j in enumerate(dims)]
        s = time.perf_counter()
        out = self.pool.map(_nm_dim, tasks, chunksize=max(1, len(tasks) // (4 * C)))
        wall_nm = time.perf_counter() - s
        nfev = int(sum(o[0] for o in out))
        cpu_nm = float(sum(o[1] for o in out))
        n_dims = len(tasks)
        t_predict = t_knn + wall_nm * (self.d / n_dims)
        t_iter = math.ceil(N / C) * t_F_slice + (N - 1) * (self.t_G + t_predict)
        return dict(t_iter=t_iter, t_F_slice=t_F_slice, t_G=self.t_G, t_predict=t_predict, t_knn=t_knn,
                    nm_runs=n_dims * 9, nfev=nfev, cpu_s_per_nm_run=cpu_nm / (n_dims * 9), wall_nm=wall_nm,
                    wall_F=wall_F, n_dims=n_dims, n_slices=n_slices, cores=C,
                    fits_per_s=(N - 1) * self.d / ((N - 1) * (self.t_G + t_predict)))
