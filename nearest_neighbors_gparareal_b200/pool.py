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
    """sharded=True (with an initialised torch.distributed NCCL group of W ranks, one per GPU): each
    rank propagates a contiguous block of the slices and the blocks are exchanged with ONE
    all-gather -- the multi-GPU analogue of the reference's MPIPoolExecutor (FHN_PDE.py:123-126)."""

    def __init__(self, sharded=False, group=None):
        self.sharded = sharded
        self.group = group

    def _batch(self, batch, t0, t1, u0):
        if self.sharded:
            import torch
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
                n, d = u0.shape
                chunk = (n + world - 1) // world
                lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
                buf = torch.zeros((world * chunk, d), dtype=torch.float64, device='cuda')
                if hi > lo:
                    buf[lo:hi] = torch.from_numpy(batch(t0[lo:hi], t1[lo:hi], u0[lo:hi])).cuda()
                dist.all_gather_into_tensor(buf, buf[rank * chunk:(rank + 1) * chunk], group=self.group)
                return buf[:n].cpu().numpy()
        return batch(t0, t1, u0)

    @staticmethod
    def _which(fn):
        """(owner, 'F'|'G', timed) when fn is run_F / run_F_timed / run_G / run_G_timed of a solver that
        can batch; identified by function identity, because the reference's timing decorator
        (solver.py:21-27) does not preserve __name__."""
        owner = getattr(fn, "__self__", None)
        func = getattr(fn, "__func__", None)
        if owner is None or func is None or not hasattr(owner, "run_F_batch"):
            return None
        for name, kind, timed in (("run_F_timed", "F", True), ("run_F", "F", False),
                                  ("run_G_timed", "G", True), ("run_G", "G", False)):
            if getattr(type(owner), name, None) is func:
                return owner, kind, timed
        return None

    def map(self, fn, *iterables, chunksize=None):
        hit = self._which(fn)
        if hit is not None:
            owner, kind, timed = hit
            t0, t1, u0 = [list(it) for it in iterables]
            if len(t0) == 0:
                return []
            s = time.time()
            batch = owner.run_F_batch if kind == "F" else owner.run_G_batch
            u1 = self._batch(batch, np.asarray(t0, dtype=float), np.asarray(t1, dtype=float), np.stack(u0))
            secs = (time.time() - s) / len(t0)
            if timed:
                return [(u1[i], secs) for i in range(len(t0))]
            return [u1[i] for i in range(len(t0))]
        return map(fn, *iterables)

    def shutdown(self, *args, **kwargs):
        pass
