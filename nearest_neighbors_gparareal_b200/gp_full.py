"""GParareal's full-dataset GP model (reference models.py:273-473, `GPjax_p`) on the GPU.

SURVEY.md section 8f rank 2: NOT part of the nnGParareal hot path -- it is the comparison model of BASELINE.json
configs[2] ("nnGParareal vs GParareal vs Parareal").  Its arithmetic is one dense n_data x n_data Cholesky per
objective evaluation, i.e. plain library linear algebra: the evaluations of all d x 9 Nelder-Mead searches of a fit
are batched and factorised with torch.linalg.cholesky_ex (cuSOLVER potrfBatched, FP64) -- a LIBRARY call, stated as
such; the hand-written kernels of this package are the nnGP path.  What is reproduced from the reference:

  * kernel sigma_y^2 exp(-r^2 / (2 sigma_x^2)) in natural (not log10) parameters     models.py:303-307
  * objective 0.5 y' K^-1 y + sum log L_ii + (n/2) log 2 pi, +inf on a failed factorisation   :317-331
  * per dimension 9 jitters 1e-20..1e-12, SciPy Nelder-Mead from the previous optimum (warm start, first [1, 1]),
    fatol = xatol = 1e-4, selection fval < 0.9 min then first minimum                    :386-421
  * random-restart fallback when every jitter fails: max(3, N/9) x 9 searches from 10^U(-4,1) drawn from
    default_rng(45)                                                                      :358-384
  * prediction with the (L, alpha) of a dimension memoised until the next fit            :434-453
  * timing keys                                                                          :296-300

The searches advance in lockstep rounds: every active search hands in the one point its SciPy state machine asks for
next, the batch is evaluated on the device, every search consumes its value.
"""
import time

import numpy as np

from .models import ModelAbstr

JITTERS = np.arange(-20, -11, dtype=float)
_HALF_LOG_2PI = 0.5 * np.log(2 * np.pi)


class _NM:
    """SciPy's non-adaptive Nelder-Mead for N = 2 as a step function (same decisions as csrc/gpfit.cu::nelder_mead)."""
    __slots__ = ("sim", "fsim", "pt", "phase", "fcalls", "it", "xbar", "xr", "fxr", "done", "fatol", "xatol")

    def __init__(self, x0, fatol, xatol):
        x0 = np.asarray(x0, dtype=float)
        self.sim = np.stack([x0, x0, x0])
        for k in range(2):
            self.sim[k + 1, k] = 1.05 * x0[k] if x0[k] != 0 else 0.00025
        self.fsim = np.full(3, np.inf)
        self.phase, self.fcalls, self.it, self.done = 0, 0, 1, False
        self.pt = self.sim[0].copy()
        self.fatol, self.xatol = fatol, xatol
        self.xbar = self.xr = None
        self.fxr = 0.0

    def _sort(self):
        order = np.argsort(self.fsim, kind="stable")
        self.sim, self.fsim = self.sim[order], self.fsim[order]

    def _next_iteration(self):
        self._sort()
        if not (self.fcalls < 400 and self.it < 400):
            self.done = True
            return
        with np.errstate(invalid="ignore"):
            if (np.max(np.abs(self.sim[1:] - self.sim[0])) <= self.xatol
                    and np.max(np.abs(self.fsim[0] - self.fsim[1:])) <= self.fatol):
                self.done = True
                return
        self.xbar = (self.sim[0] + self.sim[1]) / 2
        self.xr = 2 * self.xbar - self.sim[2]
        self.phase, self.pt = 3, self.xr

    def step(self, f):
        """consume the objective value of self.pt; sets the next point or done"""
        self.fcalls += 1
        ph = self.phase
        budget = self.fcalls >= 400  # a further evaluation would raise inside SciPy: the iteration's update is dropped
        if ph < 3:  # the three initial vertices
            self.fsim[ph] = f
            if ph < 2:
                self.phase, self.pt = ph + 1, self.sim[ph + 1].copy()
                return
            self._next_iteration()
            return
        worst = self.sim[2]
        if ph == 3:  # reflection
            self.fxr = f
            nxt = None
            if f < self.fsim[0]:
                nxt = (4, 3 * self.xbar - 2 * worst)
            elif f < self.fsim[1]:
                self.sim[2], self.fsim[2] = self.xr, f
            elif f < self.fsim[2]:
                nxt = (5, 1.5 * self.xbar - 0.5 * worst)
            else:
                nxt = (6, 0.5 * self.xbar + 0.5 * worst)
            if nxt is not None:
                if budget:
                    self._next_iteration()
                    return
                self.phase, self.pt = nxt
                return
        elif ph == 4:  # expansion
            if f < self.fxr:
                self.sim[2], self.fsim[2] = self.pt, f
            else:
                self.sim[2], self.fsim[2] = self.xr, self.fxr
        elif ph in (5, 6):  # outside / inside contraction
            if (f <= self.fxr) if ph == 5 else (f < self.fsim[2]):
                self.sim[2], self.fsim[2] = self.pt, f
            else:
                self.sim[1] = self.sim[0] + 0.5 * (self.sim[1] - self.sim[0])
                if budget:
                    self._next_iteration()
                    return
                self.phase, self.pt = 7, self.sim[1].copy()
                return
        elif ph == 7:  # first shrunk vertex
            self.fsim[1] = f
            self.sim[2] = self.sim[0] + 0.5 * (self.sim[2] - self.sim[0])
            if budget:
                self._next_iteration()
                return
            self.phase, self.pt = 8, self.sim[2].copy()
            return
        else:  # second shrunk vertex
            self.fsim[2] = f
        self.it += 1
        self._next_iteration()

    def result(self):
        return self.sim[0].copy(), float(np.min(self.fsim))


class CudaGP(ModelAbstr):
    def __init__(self, n, N, worker_pool=None, theta=None, jitter=None, fatol=None, xatol=None, device=None, **kwargs):
        super().__init__(N=N, **kwargs)
        theta = np.array([1, 1] if theta is None else theta, dtype=float)
        self.name = 'GP'
        self.hyp = np.ones((n, theta.shape[0], N))
        self.thetas = [theta for _ in range(n)]
        self.jitters = [None for _ in range(n)]
        self.fatol = 1e-4 if fatol is None else fatol
        self.xatol = 1e-4 if xatol is None else xatol
        self.theta, self.N, self.n = theta, N, n
        self.mem = {}
        self.pool = worker_pool
        self.rng = np.random.default_rng(45)
        self.tot_train_t = np.zeros(N)
        self.train_count = np.zeros(N)
        self.max_batch_bytes = kwargs.get('max_batch_bytes', 2 << 30)
        self._device = device
        self.k = 0

    def get_times(self):
        out = super().get_times()
        with np.errstate(all="ignore"):
            avg = (self.tot_train_t / self.train_count)[:self.k + 1]
        out.update({'serial_train_time': self.tot_train_t[:self.k + 1], 'avg_serial_train_time': avg})
        return out

    # ---- device evaluation ----------------------------------------------------------------------------------
    def _dev(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("CudaGP needs a CUDA device: there is no CPU fallback")
        return torch.device('cuda', torch.cuda.current_device() if self._device is None else self._device)

    def _upload(self, x, y):
        import torch
        dev = self._dev()
        self._x = torch.from_numpy(np.ascontiguousarray(x, dtype=float)).to(dev)
        self._y = torch.from_numpy(np.ascontiguousarray(y, dtype=float)).to(dev)
        # squared distances of the data, once per fit (models.py:306 recomputes them in every evaluation); formed as
        # sum_j (x_ij - x_kj)^2 like cdist('sqeuclidean'), in row blocks
        n, d = self._x.shape
        self._r2 = torch.empty((n, n), dtype=torch.float64, device=dev)
        blk = max(1, min(n, (256 << 20) // max(1, 8 * n * d)))
        for lo in range(0, n, blk):
            diff = self._x[lo:lo + blk, None, :] - self._x[None, :, :]
            self._r2[lo:lo + blk] = (diff * diff).sum(-1)
        self._eye = torch.eye(x.shape[0], dtype=torch.float64, device=dev)

    def _nll_batch(self, thetas, jitters, dims):
        """negative log marginal likelihood for a batch of (theta, jitter, output dimension); +inf on failure"""
        import torch
        n = self._r2.shape[0]
        out = np.empty(len(dims))
        chunk = max(1, int(self.max_batch_bytes // (8 * n * n * 3)))
        th = torch.from_numpy(np.asarray(thetas, dtype=float)).to(self._r2.device)
        jt = torch.from_numpy(10.0 ** np.asarray(jitters, dtype=float)).to(self._r2.device)
        dm = torch.from_numpy(np.asarray(dims, dtype=np.int64)).to(self._r2.device)
        for lo in range(0, len(dims), chunk):
            sl = slice(lo, lo + chunk)
            sx, sy = th[sl, 0], th[sl, 1]
            K = (sy * sy)[:, None, None] * torch.exp((-0.5 * (1 / (sx * sx)))[:, None, None] * self._r2[None])
            K = K + jt[sl, None, None] * self._eye[None]
            L, info = torch.linalg.cholesky_ex(K)
            yb = self._y[:, dm[sl]].T.unsqueeze(-1)                      # [b, n, 1]
            z = torch.linalg.solve_triangular(L, yb, upper=False)
            quad = (z * z).sum((1, 2))
            logdet = torch.log(torch.diagonal(L, dim1=1, dim2=2)).sum(1)
            res = 0.5 * quad + logdet + n * _HALF_LOG_2PI
            res = torch.where((info == 0) & torch.isfinite(res), res, torch.full_like(res, float('inf')))
            out[sl] = res.cpu().numpy()
        return out

    def _search(self, x0s, jitters, dims):
        """lockstep Nelder-Mead searches: one per (start, jitter, dimension); returns thetas [S, 2], fvals [S]"""
        nms = [_NM(x0, self.fatol, self.xatol) for x0 in x0s]
        active = list(range(len(nms)))
        while active:
            pts = [nms[s].pt for s in active]
            f = self._nll_batch(pts, [jitters[s] for s in active], [dims[s] for s in active])
            nxt = []
            for s, fv in zip(active, f):
                nms[s].step(float(fv))
                if not nms[s].done:
                    nxt.append(s)
            active = nxt
        res = [nm.result() for nm in nms]
        return np.array([r[0] for r in res]), np.array([r[1] for r in res]), sum(nm.fcalls for nm in nms)

    @staticmethod
    def _select(fvals):
        """models.py:399-402: mask fval < 0.9 * min, all if empty, first minimum"""
        fvals = np.asarray(fvals, dtype=float)
        with np.errstate(all="ignore"):
            mask = fvals < fvals.min() * 0.9
        if mask.sum() == 0:
            mask[:] = True
        idx = np.flatnonzero(mask)
        return int(idx[int(np.argmin(fvals[idx]))])

    def _train_coord_rnd(self, coord):
        """models.py:358-384: random restarts for one output dimension"""
        tot_rnd = max(3, int(self.N / 9))
        jit = np.tile(JITTERS, tot_rnd)
        thetas = [10 ** self.rng.uniform(-4, 1, 2) for _ in range(len(jit))]
        s = time.time()
        th, fv, _ = self._search(thetas, jit, [coord] * len(jit))
        self.tot_train_t[self.k] += time.time() - s
        self.train_count[self.k] += len(jit)
        b = self._select(fv)
        if np.isinf(fv[b]):
            print('random restart failed')
            return self._train_coord_rnd(coord)
        return th[b], fv[b], jit[b]

    # ---- reference protocol ---------------------------------------------------------------------------------
    def fit(self, x, y, k, *args, **kwargs):
        """models.py:424-431 + :386-421"""
        self.mem = {}
        self.k = k
        self.x, self.y = x, y
        self._upload(x, y)
        n = self.n
        dims = np.repeat(np.arange(n), len(JITTERS))
        jit = np.tile(JITTERS, n)
        x0s = [np.asarray(self.thetas[j], dtype=float) for j in dims]
        s = time.time()
        th, fv, nfev = self._search(x0s, jit, dims)
        self.tot_train_t[k] += time.time() - s
        self.train_count[k] += len(dims)
        self.nfev_last = nfev
        new = np.zeros((n, 2))
        for j in range(n):
            sl = slice(j * len(JITTERS), (j + 1) * len(JITTERS))
            b = self._select(fv[sl])
            opt, fval, jt = th[sl][b], fv[sl][b], JITTERS[b]
            if np.isinf(fval):
                print('------> GP trainign failed for coordinate', j)
                opt, fval, jt = self._train_coord_rnd(j)
            self.thetas[j], self.jitters[j] = tuple(opt), jt
            new[j] = opt
        self.hyp[..., k + 1] = new
        self._alpha = None

    def _prepare(self):
        """(L, alpha) of every dimension at its optimum (the memo of models.py:434-445), batched"""
        import torch
        n = self._r2.shape[0]
        th = torch.from_numpy(np.array(self.thetas, dtype=float)).to(self._r2.device)
        jt = torch.from_numpy(10.0 ** np.array(self.jitters, dtype=float)).to(self._r2.device)
        alpha = torch.empty((self.n, n), dtype=torch.float64, device=self._r2.device)
        chunk = max(1, int(self.max_batch_bytes // (8 * n * n * 3)))
        for lo in range(0, self.n, chunk):
            sl = slice(lo, lo + chunk)
            sx, sy = th[sl, 0], th[sl, 1]
            K = (sy * sy)[:, None, None] * torch.exp((-0.5 * (1 / (sx * sx)))[:, None, None] * self._r2[None])
            K = K + jt[sl, None, None] * self._eye[None]
            L = torch.linalg.cholesky_ex(K)[0]
            alpha[sl] = torch.cholesky_solve(self._y[:, sl].T.unsqueeze(-1), L).squeeze(-1)
        self._alpha, self._th = alpha, th

    def predict(self, new_x, prev_F=None, prev_G=None, *args, **kwargs):
        """models.py:448-453: post_mean_j = k_j(x, new_x)' alpha_j"""
        import torch
        if self._alpha is None:
            self._prepare()
        q = torch.from_numpy(np.asarray(new_x, dtype=float).reshape(1, -1)).to(self._r2.device)
        diff = self._x - q
        r2 = (diff * diff).sum(1)                                          # [n]
        sx, sy = self._th[:, 0:1], self._th[:, 1:2]
        ks = (sy * sy) * torch.exp((-0.5 * (1 / (sx * sx))) * r2[None])    # [d, n]
        return (ks * self._alpha).sum(1).cpu().numpy()

    def store(self):
        keep = {k: getattr(self, k, None) for k in ('_x', '_y', '_r2', '_eye', '_alpha', '_th', 'mem', 'pool')}
        for k in keep:
            setattr(self, k, None)
        try:
            new = super().store()
        finally:
            for k, v in keep.items():
                setattr(self, k, v)
        new.hyp = new.hyp[..., :self.k + 3]
        return new

    def restore_attrs(self, pool):
        self.pool = pool
        self.mem = {}
        hyp = np.ones((self.n, self.theta.shape[0], self.N))
        hyp[..., :self.hyp.shape[-1]] = self.hyp
        self.hyp = hyp
        if getattr(self, 'x', None) is not None:
            self._upload(self.x, self.y)
            self._alpha = None


GPjax_p = CudaGP
