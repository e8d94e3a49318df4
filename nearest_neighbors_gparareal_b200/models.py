"""Host mirror of the reference's model protocol (models.py:19-270).

`CudaNNGP` keeps the constructor, kwargs (`nn`, `n_restarts`, `seed`, `fatol`, `xatol`, `theta`,
`calc_detail_avg`, `calc_parall_overhead`), methods (`fit`, `predict`, `fit_timed`,
`predict_timed`, `get_times`, `store`, `restore_attrs`) and timing keys of `NNGP_p`.  The kNN,
the d*9*R Nelder-Mead fits, the selection and the posterior mean run in csrc/knn.cu and
csrc/gpfit.cu through ONE C-ABI call per predict (nngp_predict_host).  What stays on the host
is what the reference keeps on the host: the NumPy `default_rng(seed)` stream of Nelder-Mead
start points (models.py:114,192), drawn in the same order.
"""
import copy
import time

import numpy as np

from . import _lib
from .utils import dim_block

N_JITTER = 9  # models.py:186  jitter = np.arange(-20, -11)


class ModelAbstr():
    def __init__(self, **kwargs):
        self.train_time = 0
        self.pred_time = 0
        N = kwargs['N']
        self.pred_times = np.zeros(N)

    def fit_timed(self, x, y, *args, **kwargs):
        self.time_k = kwargs['k']
        s_time = time.time()
        ret = self.fit(x, y, *args, **kwargs)
        elap_time = time.time() - s_time
        self.train_time += elap_time
        self.pred_times[self.time_k] += elap_time
        return ret

    def predict_timed(self, new_x, *args, **kwargs):
        s_time = time.time()
        ret = self.predict(new_x, *args, **kwargs)
        elap_time = time.time() - s_time
        self.pred_time += elap_time
        self.pred_times[self.time_k] += elap_time
        return ret

    def get_times(self):
        return {'mdl_train_t': self.train_time, 'mdl_pred_t': self.pred_time,
                'mdl_tot_t': self.train_time + self.pred_time,
                'by_iter': self.pred_times[:getattr(self, 'time_k', -1) + 1]}

    def fit(self, x, y, *args, **kwargs):
        raise Exception('Not implemented')

    def predict(self, new_x, prev_F, prev_G):
        raise Exception('Not implemented')

    def restore_attrs(self, pool):
        """re-attach what store() stripped (models.py:262-270); nothing to do for models without a pool"""
        pass

    def store(self):
        saved = {k: self.__dict__.get(k) for k in ('pool', '_handle') if k in self.__dict__}
        for k in saved:
            self.__dict__[k] = None
        new = copy.deepcopy(self)
        self.__dict__.update(saved)
        return new


class BareParareal(ModelAbstr):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.name = 'Parareal'

    def fit(self, *args, **kwargs):
        pass

    def predict(self, new_x, prev_F, prev_G, *args, **kwargs):
        return prev_F - prev_G


class CudaNNGP(ModelAbstr):
    def __init__(self, n, N, worker_pool=None, theta=None, fatol=None, xatol=None, handle=None, **kwargs):
        super().__init__(N=N, **kwargs)
        if theta is None:
            theta = [1, 1]
        self.theta = np.array(theta)
        if self.theta.shape[0] != 2:
            raise Exception('the squared-exponential kernel has exactly two hyper-parameters')
        self.name = 'NNGP'
        self.fatol = 1e-1 if fatol is None else fatol
        self.xatol = 1e-1 if xatol is None else xatol
        self.n = n
        self.N = N
        self.n_restarts = kwargs.get('n_restarts', 1)
        self.nn = kwargs.get('nn', 'adaptive')
        self.seed = kwargs.get('seed', 45)
        self.rng = np.random.default_rng(self.seed)
        np.random.seed(self.seed)
        self.pool = worker_pool  # unused: the GPU is the worker pool
        self.tot_train_t = 0
        self.train_count = 0
        self.nfev_total = 0
        self.calc_detail_avg = kwargs.get('calc_detail_avg', False)
        self.calc_parall_overhead = kwargs.get('calc_parall_overhead', False)
        self.collect_nfev = kwargs.get('collect_nfev', False)
        # W > 1 torch.distributed ranks calling predict in lockstep (the host loop under torchrun): each rank fits
        # d/W of the output dimensions and the predictions are all-gathered (needs d % W == 0)
        self.shard_predict = kwargs.get('shard_predict', False)
        self.group = kwargs.get('group', None)
        if self.calc_detail_avg:
            self.detail_avg = np.zeros((N, N))
        if self.calc_parall_overhead:
            self.overhead = np.zeros((N, N))
        self._handle = handle
        self._n_dev = 0          # dataset rows already on the device
        self.k = 0

    # -- helpers -------------------------------------------------------------------------
    def handle(self):
        if self._handle is None:
            self._handle = _lib.default_handle()
        return self._handle

    def neighbours(self, k=None):
        """models.py:172-175"""
        k = self.k if k is None else k
        return max(10, k + 2) if self.nn == 'adaptive' else self.nn

    def draw_starts(self, n_predicts=1):
        """models.py:190-192: one rng.integers(-8, 0, 2) per (dim, jitter, restart) task, in task
        order; the vectorised draw consumes the PCG64 stream identically (tests/test_host.py)."""
        n_tasks = self.n * N_JITTER * self.n_restarts
        raw = self.rng.integers(-8, 0, (n_predicts * n_tasks, 2))
        return raw.astype(np.int8).reshape(n_predicts, self.n, N_JITTER, self.n_restarts, 2)

    # -- reference protocol --------------------------------------------------------------
    def fit(self, x, y, k, *args, **kwargs):
        """models.py:157-159 stores x, y, k; here the device dataset is brought up to date with
        the rows appended since the last call (parareal.py:336-339 only ever appends)."""
        self.k = k
        self.x, self.y = x, y
        h = self.handle()
        rows = x.shape[0]
        if h.dataset_dim() != x.shape[1] or h.dataset_rows() != self._n_dev or rows < self._n_dev:
            h.dataset_reset()
            h.dataset_reserve(max(rows * 2, 1024), x.shape[1])
            self._n_dev = 0
        if rows > self._n_dev:
            h.dataset_append_host(x[self._n_dev:], y[self._n_dev:])
            self._n_dev = rows

    def predict(self, new_x, prev_F=None, prev_G=None, *args, **kwargs):
        """models.py:171-226"""
        details = kwargs.get('return_details', False)
        h = self.handle()
        m = min(self.neighbours(), self._n_dev)
        if m > 160:
            raise Exception('nn > 160 neighbours is not supported (include/nngpara.h: NNGP_MAX_NEIGHBOURS_BIG)')
        new_x = np.asarray(new_x, dtype=float).reshape(1, -1)
        starts = self.draw_starts(1)
        s = time.time()
        block = self._block()
        dev_buf = self._device_gather_buffer() if block is not None else None
        out = h.predict_host(new_x, m, starts, self.n_restarts, self.fatol, self.xatol,
                             details=details or self.collect_nfev, block=block, pred_out=dev_buf)
        if dev_buf is not None:
            # the rank's block is already on the device: all-gather in place, one read back
            import torch.distributed as dist
            j0, dl = block
            dist.all_gather_into_tensor(dev_buf, dev_buf[j0:j0 + dl], group=self.group)
            out['pred'] = dev_buf.cpu().numpy()[None, :]
        elif block is not None:
            out['pred'] = self._gather(out['pred'], block)
        el = time.time() - s
        n_tasks = self.n * N_JITTER * self.n_restarts
        self.tot_train_t += el
        self.train_count += n_tasks
        if self.collect_nfev:
            self.nfev_total += int(out['nfev'].sum())
        i = kwargs.get('i', None)
        if self.calc_detail_avg and i is not None:
            self.detail_avg[self.k, i] = el / n_tasks
        if self.calc_parall_overhead and i is not None:
            self.overhead[self.k, i] = 0.0
        preds = out['pred'][0]
        return (preds, out) if details else preds

    def _block(self):
        if not self.shard_predict:
            return None
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(self.group)
        if world == 1 or self.n % world != 0:
            return None
        return dim_block(self.n, dist.get_rank(self.group), world)

    def _device_gather_buffer(self):
        """a d-double device tensor on the handle's GPU when the group's backend is NCCL (else None: host gather)"""
        import torch
        import torch.distributed as dist
        if dist.get_backend(self.group) != 'nccl':
            return None
        if getattr(self, '_gather_buf', None) is None:
            self._gather_buf = torch.empty(self.n, dtype=torch.float64, device=torch.device('cuda', self.handle().device))
        return self._gather_buf

    def _gather(self, pred, block):
        """all-gather of the ranks' prediction blocks (d/W doubles each); details stay per-rank"""
        import torch
        import torch.distributed as dist
        j0, dl = block
        dev = 'cuda' if dist.get_backend(self.group) == 'nccl' else 'cpu'
        t = torch.from_numpy(np.ascontiguousarray(pred[0])).to(dev)
        mine = t[j0:j0 + dl] if dev == 'cuda' else t[j0:j0 + dl].clone()
        dist.all_gather_into_tensor(t, mine, group=self.group)
        return t.cpu().numpy()[None, :]

    def get_times(self):
        out = super().get_times()
        detail_avg = self.detail_avg[:self.k + 1, :] if self.calc_detail_avg else None
        overhead = self.overhead[:self.k + 1, :] if self.calc_parall_overhead else None
        out.update({'serial_train_time': self.tot_train_t, 'calc_detail_avg': detail_avg, 'overhead': overhead,
                    'avg_serial_train_time': self.tot_train_t / max(self.train_count, 1)})
        return out

    def store(self):
        new = super().store()
        new.pool = None
        return new

    def restore_attrs(self, pool):
        self.pool = pool
        self._handle = None
        self._n_dev = 0


NNGP_p = CudaNNGP
