"""Host mirror of the reference driver (parareal.py:26-471, 812-1060).

`Parareal` / `PararealLight` run the reference's outer loop on the host against the solver /
model / executor protocols, so any `SolverAbstr` / `ModelAbstr` plugs in; with `CudaSolverRK`,
`CudaNNGP` and `CudaPool` every arithmetic step is a C-ABI call on host buffers (this is the
end-to-end path bench.py reports as `e2e`).

`PararealDevice` keeps the whole iteration resident on the GPU: one batched RK launch for the
fine solves (sharded by time slice across ranks, exchanged with a single all-gather), a device
dataset append, the fused on-device sweep (nngp_sweep) and one small device->host read of the
per-slice errors per iteration.
"""
import time

import numpy as np

from . import _lib
from .models import BareParareal, CudaNNGP, ModelAbstr, N_JITTER
from .pool import CudaPool, MyPool
from .solver import CudaSolverRK, SolverAbstr
from .systems import ODE


class Parareal():
    def __init__(self, ode, solver, tspan, N, epsilon=5e-7, verbose='v', **kwargs):
        if not isinstance(ode, ODE):
            raise Exception('ode must be an instance of the ODE class, see systems.py file.')
        if not isinstance(solver, SolverAbstr):
            raise Exception('solver must be an instance of the SolverAbstr class, see solver.py file.')
        self.tspan = tspan
        self.N = N
        self.epsilon = epsilon
        self.runs = dict()
        self.fine = None
        self.ode_name = ode.name
        self.n = ode.get_dim()
        self.ode = ode
        self.solver = solver
        self.f = ode.get_vector_field()
        self.u0 = ode.get_init_cond()
        self.verbose = verbose

    keep_history = True  # Parareal returns u[N+1, d, K]; PararealLight only the last iterate

    def _get_pool(self, *args, **kwargs):
        pool = kwargs.get('pool', None)
        if isinstance(pool, int) or pool is None:
            # the reference spawns `pool` CPU worker processes (parareal.py:60-61); here the GPU
            # is the worker pool whatever the requested count
            pool = CudaPool()
        return pool

    def run(self, *args, **kwargs):
        pool = self._get_pool(*args, **kwargs)
        kwargs['pool'] = pool
        try:
            out = self._run(*args, **kwargs)
        except Exception:
            pool.shutdown()
            raise
        pool.shutdown()
        return out

    def _make_model(self, model, **kwargs):
        if isinstance(model, ModelAbstr):
            return model
        name = model.lower()
        if name == 'parareal':
            return BareParareal(N=self.N, **kwargs)
        if name == 'nngp':
            if 'pool' not in kwargs:
                raise Exception('A worker pool must be provided to run NNGP in parallel')
            kw = dict(kwargs)
            pool = kw.pop('pool')
            return CudaNNGP(n=self.n, N=self.N, worker_pool=pool, **kw)
        # 'gpjax' (full GParareal) and 'elm' are outside the nnGParareal hot path
        raise Exception('Not implemented')

    def _run(self, model='parareal', cstm_mdl_name=None, add_model=False, **kwargs):
        mdl = self._make_model(model, **kwargs)
        s_time = time.time()
        out = self._parareal(mdl, **kwargs)
        elap_time = time.time() - s_time
        out['timings']['runtime'] = elap_time
        if self.verbose == 'v':
            print(f'Elapsed Parareal time: {elap_time:0.2f}s')
        if add_model:
            out['mdl'] = mdl.store()
        if cstm_mdl_name is None:
            cstm_mdl_name = mdl.name
        self.runs[cstm_mdl_name] = out
        return out

    # the iteration of parareal.py:212-471, state kept as in PararealLight (:851-862)
    def _parareal(self, model, debug=False, early_stop=None, parall='Serial', store_int=False, **kwargs):
        if store_int:
            raise NotImplementedError('intermediate checkpoints are outside the hot path (SURVEY.md section 8f)')
        N, eps, n = self.N, self.epsilon, self.n
        solver = self.solver
        t = np.linspace(self.tspan[0], self.tspan[1], num=N + 1)
        parall = parall.lower()
        pool = kwargs.get('pool', None)
        if parall == 'mpi' and pool is None:
            raise Exception('MPI parallel backend requested but no pool of worker provided')
        verbose = kwargs.get('verbose', self.verbose)
        I = 0
        conv_int = []
        err = np.full((N + 1, N), np.nan)
        cur = {key: np.full((N + 1, n), np.nan) for key in ('u', 'uG', 'uF')}
        for a in cur.values():
            a[0] = self.u0
        G_time = F_time = F_time_serial = 0
        # coarse initialisation (parareal.py:264-277)
        state = self.u0
        for i in range(N):
            state, secs = solver.run_G_timed(t[i], t[i + 1], state)
            G_time += secs
            cur['uG'][i + 1] = state
        cur['u'][:] = cur['uG']
        nxt = {key: a.copy() for key, a in cur.items()}
        history = [cur['u'].copy()]
        x = np.zeros((0, n))
        D = np.zeros((0, n))
        k = 0
        for k in range(N):
            if verbose == 'v':
                print(f'{self.ode_name} {model.name} iteration number (out of {N}): {k+1} ')
            s_time = time.time()
            # fine solves of the unconverged slices (parareal.py:309-329)
            if parall in ('mpi', 'joblib'):
                res = list(pool.map(solver.run_F_timed, t[I:N], t[I + 1:N + 1], [cur['u'][i] for i in range(I, N)]))
                cur['uF'][I + 1:N + 1] = np.array([r[0] for r in res])
                F_time_serial += np.array([r[1] for r in res]).mean()
            else:
                last = 0
                for i in range(I, N):
                    cur['uF'][i + 1], last = solver.run_F_timed(t[i], t[i + 1], cur['u'][i])
                F_time_serial += last / (N - I)
            F_time += time.time() - s_time
            # slice I+1 is exact now (parareal.py:331-334)
            nxt['uG'][I + 1] = cur['uG'][I + 1]
            nxt['uF'][I + 1] = cur['uF'][I + 1]
            nxt['u'][I + 1] = cur['uF'][I + 1]
            I += 1
            # training data (parareal.py:336-339)
            x = np.vstack([x, cur['u'][I - 1:N]])
            D = np.vstack([D, cur['uF'][I:N + 1] - cur['uG'][I:N + 1]])
            if I == N:
                if verbose == 'v':
                    print('WARNING: early stopping')
                err[:, k] = np.linalg.norm(nxt['u'] - cur['u'], np.inf, 1)
                err[-1, k] = np.nextafter(eps, 0)
                history.append(nxt['u'].copy())
                break
            model.fit_timed(x, D, k=k)
            # serial sweep (parareal.py:359-382)
            for i in range(I, N):
                nxt['uG'][i + 1], secs = solver.run_G_timed(t[i], t[i + 1], nxt['u'][i])
                G_time += secs
                preds = model.predict_timed(nxt['u'][i].reshape(1, -1), cur['uF'][i + 1], cur['uG'][i + 1], i=i)
                nxt['u'][i + 1] = preds + nxt['uG'][i + 1]
            if np.any(np.isnan(nxt['uG'])):
                raise Exception("NaN values in initial coarse solve - increase Ng!")
            # convergence bookkeeping (parareal.py:402-416)
            err[:, k] = np.linalg.norm(nxt['u'] - cur['u'], np.inf, 1)
            err[I, k] = 0
            cur['u'][:] = nxt['u']
            cur['uG'][:] = nxt['uG']
            for p in range(I + 1, N + 1):
                if err[p, k] < eps:
                    nxt['uF'][p] = cur['uF'][p]
                    I += 1
                else:
                    break
            cur['uF'][:] = nxt['uF']
            history.append(cur['u'].copy())
            if verbose == 'v':
                print('--> Converged:', I)
            conv_int.append(I)
            if I == N:
                break
            if (early_stop is not None) and k == (early_stop - 1):
                if verbose == 'v':
                    print('Early stopping due to user condition.')
                break
        timings = {'F_time': F_time, 'G_time': G_time, 'F_time_serial_avg': F_time_serial}
        timings.update(model.get_times())
        u_out = np.stack(history[:k + 1], axis=2) if self.keep_history else history[-1]
        return {'t': t, 'u': u_out, 'u_last': history[-1], 'err': err[:, :k + 1], 'x': x, 'D': D, 'k': k + 1,
                'timings': timings, 'debug_dict': {}, 'converged': I == N, 'conv_int': conv_int}


class PararealLight(Parareal):
    """parareal.py:782-1060: same loop, returns only the last iterate u[N+1, d]."""
    keep_history = False


class PararealDevice(Parareal):
    """Device-resident nnGParareal: the whole iteration of parareal.py:301-439 on the GPU.

    Needs a CudaSolverRK and either the 'nngp' or the 'parareal' model.  With an initialised
    torch.distributed process group of W>1 ranks (one per GPU) the fine solves are sharded by
    time slice: rank r propagates a contiguous block of the unconverged slices, writing straight
    into its part of the gather buffer, and ONE all-gather per iteration makes every rank hold
    all uF rows; the serial sweep is replicated on every rank (bit-identical results).
    """
    keep_history = False

    def __init__(self, ode, solver, tspan, N, epsilon=5e-7, verbose='v', group=None, **kwargs):
        super().__init__(ode, solver, tspan, N, epsilon=epsilon, verbose=verbose, **kwargs)
        if not isinstance(solver, CudaSolverRK):
            raise Exception('PararealDevice needs a CudaSolverRK')
        self.group = group
        self.events = []

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def _parareal(self, model, early_stop=None, parall='Serial', store_int=False, return_data=True,
                  max_rows=None, **kwargs):
        import torch
        if store_int:
            raise NotImplementedError('intermediate checkpoints are outside the hot path')
        if not isinstance(model, (CudaNNGP, BareParareal)):
            raise Exception('PararealDevice supports the nngp and parareal models')
        N, eps, n = self.N, self.epsilon, self.n
        solver = self.solver
        h, sysid = solver.device()
        dev = torch.device('cuda', h.device)
        rank, world = self._world()
        verbose = kwargs.get('verbose', self.verbose)
        mF, mG = _lib.METHODS[solver.F], _lib.METHODS[solver.G]
        if max(solver.Nf, solver.Ng) > solver.thresh:
            raise Exception('steps > thresh (RK paging, solver.py:89-96) is only supported by the host driver')
        stream = torch.cuda.current_stream(dev).cuda_stream
        f64 = dict(dtype=torch.float64, device=dev)
        t_host = np.linspace(self.tspan[0], self.tspan[1], num=N + 1)
        d_t = torch.from_numpy(t_host).to(dev)
        pad = world  # the in-place all-gather view may run past row N
        u_cur = torch.zeros((N + 1, n), **f64)
        uG_cur = torch.zeros((N + 1, n), **f64)
        uF = torch.zeros((N + 1 + N + pad, n), **f64)
        u0 = torch.from_numpy(self.u0).to(dev)
        u_cur[0] = u0
        uG_cur[0] = u0
        uF[0] = u0
        d_err = torch.zeros(N + 1, **f64)
        is_gp = isinstance(model, CudaNNGP)
        if is_gp:
            model._handle = h
            h.dataset_reset()
            h.dataset_reserve(max_rows or min(N * (N + 3) // 2 + 1, N * 64), n)
            model._n_dev = 0
        G_time = F_time = sweep_time = 0.0
        tic = time.time()
        # coarse initialisation (parareal.py:264-277): N dependent one-slice launches
        for i in range(N):
            h.rk_batch(sysid, mG, solver.h_mode, solver.Ng, 1, d_t[i:], d_t[i + 1:], uG_cur[i], n, uG_cur[i + 1], n, stream)
        u_cur.copy_(uG_cur)
        u_next = u_cur.clone()
        uG_next = uG_cur.clone()
        torch.cuda.synchronize(dev)
        G_time += time.time() - tic
        I = 0
        conv_int = []
        err = np.full((N + 1, N), np.nan)
        k = 0
        for k in range(N):
            if verbose == 'v' and rank == 0:
                print(f'{self.ode_name} {model.name} iteration number (out of {N}): {k+1} ')
            tic = time.time()
            n_act = N - I
            if world == 1:
                h.rk_batch(sysid, mF, solver.h_mode, solver.Nf, n_act, d_t[I:], d_t[I + 1:], u_cur[I], n, uF[I + 1], n, stream)
            else:
                import torch.distributed as dist
                chunk = (n_act + world - 1) // world
                lo = min(I + rank * chunk, N)
                cnt = max(0, min(chunk, N - lo))
                if cnt > 0:
                    h.rk_batch(sysid, mF, solver.h_mode, solver.Nf, cnt, d_t[lo:], d_t[lo + 1:], u_cur[lo], n, uF[lo + 1], n, stream)
                view = uF[I + 1:I + 1 + world * chunk]
                dist.all_gather_into_tensor(view, view[rank * chunk:(rank + 1) * chunk], group=self.group)
            if self.events is not None:
                torch.cuda.synchronize(dev)
                F_time += time.time() - tic
            # slice I+1 is exact (parareal.py:331-334)
            u_next[I + 1].copy_(uF[I + 1])
            uG_next[I + 1].copy_(uG_cur[I + 1])
            I += 1
            if is_gp:
                h.append_iteration(u_cur, uF, uG_cur, N, I, n, stream)
                model._n_dev = h.dataset_rows()
            if I == N:
                h.rowwise_maxabs_diff(u_next, u_cur, N + 1, n, d_err, stream)
                err[:, k] = d_err.cpu().numpy()
                err[-1, k] = np.nextafter(eps, 0)
                u_cur.copy_(u_next)
                break
            tic = time.time()
            if is_gp:
                model.k = k
                model.time_k = k
                m = min(model.neighbours(k), h.dataset_rows())
                starts = torch.from_numpy(model.draw_starts(N - I)).to(dev)
                h.sweep(sysid, mG, solver.h_mode, solver.Ng, d_t, N, I, m, model.n_restarts, starts,
                        model.fatol, model.xatol, u_next, uG_next, n, stream)
                model.train_count += (N - I) * n * N_JITTER * model.n_restarts
            else:
                for i in range(I, N):
                    h.rk_batch(sysid, mG, solver.h_mode, solver.Ng, 1, d_t[i:], d_t[i + 1:], u_next[i], n, uG_next[i + 1], n, stream)
                    torch.add(uF[i + 1] - uG_cur[i + 1], uG_next[i + 1], out=u_next[i + 1])
            # convergence bookkeeping (parareal.py:396-416): one device->host read per iteration
            h.rowwise_maxabs_diff(u_next, u_cur, N + 1, n, d_err, stream)
            err[:, k] = d_err.cpu().numpy()
            dt_sweep = time.time() - tic
            sweep_time += dt_sweep
            if is_gp:
                model.pred_time += dt_sweep
                model.pred_times[k] += dt_sweep
                model.tot_train_t += dt_sweep
            if np.any(np.isnan(err[:, k])) and bool(torch.isnan(uG_next).any()):
                raise Exception("NaN values in initial coarse solve - increase Ng!")
            err[I, k] = 0
            u_cur.copy_(u_next)
            uG_cur.copy_(uG_next)
            for p in range(I + 1, N + 1):
                if err[p, k] < eps:
                    I += 1
                else:
                    break
            if verbose == 'v' and rank == 0:
                print('--> Converged:', I)
            conv_int.append(I)
            if I == N:
                break
            if (early_stop is not None) and k == (early_stop - 1):
                break
        timings = {'F_time': F_time, 'G_time': G_time, 'F_time_serial_avg': F_time / max(N, 1), 'sweep_time': sweep_time}
        timings.update(model.get_times())
        out = {'t': t_host, 'u': u_cur.cpu().numpy(), 'err': err[:, :k + 1], 'k': k + 1, 'timings': timings,
               'debug_dict': {}, 'converged': I == N, 'conv_int': conv_int}
        out['u_last'] = out['u']
        if is_gp and return_data:
            rows = h.dataset_rows()
            out['n_rows'] = rows
        return out
