"""Executor protocol of the reference (`map`, `shutdown`; parareal.py:16-24, 58-64).

`CudaPool.map(solver.run_F_timed, t0s, t1s, u0s)` -- the unmodified call of
parareal.py:311 -- becomes ONE batched RK launch over all slices when the bound method
belongs to a CudaSolverRK; any other callable is mapped serially like `MyPool`."""
import time

import numpy as np


class MyPool():
    @staticmethod
    def map(*args, chunksize=None, **kwargs):
        return map(*args, **kwargs)

    @staticmethod
    def shutdown(*args, **kwargs):
        pass


class CudaPool():
    def map(self, fn, *iterables, chunksize=None):
        owner = getattr(fn, "__self__", None)
        name = getattr(fn, "__name__", "")
        if owner is not None and hasattr(owner, "run_F_batch") and name in (
                "run_F_timed", "run_F", "run_G_timed", "run_G"):
            t0, t1, u0 = [list(it) for it in iterables]
            if len(t0) == 0:
                return []
            s = time.time()
            batch = owner.run_F_batch if "F" in name else owner.run_G_batch
            u1 = batch(np.asarray(t0, dtype=float), np.asarray(t1, dtype=float), np.stack(u0))
            secs = (time.time() - s) / len(t0)
            if name.endswith("_timed"):
                return [(u1[i], secs) for i in range(len(t0))]
            return [u1[i] for i in range(len(t0))]
        return map(fn, *iterables)

    def shutdown(self, *args, **kwargs):
        pass
