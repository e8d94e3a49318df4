"""Intermediate checkpoints in the REFERENCE's layout (parareal.py:114-209, 420-431; models.py:64-72, 262-270).

The reference pickles the driver object with two extra attributes: `mdl` (the model's `store()` copy, RNG state
included) and `objs`, a dict with
    t, I, verbose, u / uG / uF [N+1, n, k+2], err [N+1, k+2], x, D, data_x / data_D [N, n, k+2],
    G_time, F_time, debug, early_stop, parall, store_int, kwargs, k, conv_int.
The drivers of this package keep only the current iterate; when `store_int=True` they record the iterates and write
exactly that dict, so a dump can be resumed by either side: `Parareal.load_int_dump` here accepts dumps written by
the reference (also when the reference's classes are not importable: `load_dump` falls back to attribute-bag
stand-ins) and the reference's `load_int_dump` reads the key set it expects from dumps written here.
"""
import pickle

import numpy as np

OBJ_KEYS = ('t', 'I', 'verbose', 'u', 'uG', 'uF', 'err', 'x', 'D', 'data_x', 'data_D', 'G_time', 'F_time', 'debug',
            'early_stop', 'parall', 'store_int', 'kwargs', 'k', 'conv_int')


class RunHistory:
    """records what the reference keeps in its (N+1, n, N+1) arrays, iteration by iteration"""

    def __init__(self, N, n, u0, uG0):
        self.N, self.n = N, n
        self.u = [np.array(u0, dtype=float)]      # u^0 = coarse initialisation
        self.uG = [np.array(uG0, dtype=float)]
        self.uF = []                              # uF^k: fine values computed in iteration k (+ forward-filled rows)
        self.rows = []                            # (I after the increment, appended x rows, appended D rows)

    def add_fine(self, uF_k, I, x_rows, D_rows):
        self.uF.append(np.array(uF_k, dtype=float))
        self.rows.append((int(I), np.array(x_rows, dtype=float), np.array(D_rows, dtype=float)))

    def add_iterate(self, u_next, uG_next):
        self.u.append(np.array(u_next, dtype=float))
        self.uG.append(np.array(uG_next, dtype=float))

    def arrays(self, k, I):
        """u, uG, uF [N+1, n, k+2] and data_x, data_D [N, n, k+2] after iteration k (parareal.py:425-427)"""
        N, n = self.N, self.n
        u = np.stack(self.u[:k + 2], axis=2)
        uG = np.stack(self.uG[:k + 2], axis=2)
        uF = np.full((N + 1, n, k + 2), np.nan)
        for j in range(k + 1):
            uF[:, :, j] = self.uF[j]
        uF[:I + 1, :, k + 1] = self.uF[k][:I + 1]   # converged slices are filled forward (parareal.py:331-333, 408-413)
        data_x = np.full((N, n, k + 2), np.nan)
        data_D = np.full((N, n, k + 2), np.nan)
        for j, (Ij, xr, Dr) in enumerate(self.rows[:k + 1]):
            data_x[Ij - 1:N, :, j] = xr
            data_D[Ij - 1:N, :, j] = Dr
        return u, uG, uF, data_x, data_D

    @classmethod
    def from_objs(cls, objs):
        """rebuilds the recorder from a dump (so that a resumed run can keep writing checkpoints)"""
        u, uG, uF = np.asarray(objs['u']), np.asarray(objs['uG']), np.asarray(objs['uF'])
        N, n, k = u.shape[0] - 1, u.shape[1], int(objs['k'])
        h = cls(N, n, u[:, :, 0], uG[:, :, 0])
        h.u = [u[:, :, j].copy() for j in range(k + 2)]
        h.uG = [uG[:, :, j].copy() for j in range(k + 2)]
        h.uF = [uF[:, :, j].copy() for j in range(k + 1)]
        dx, dD = np.asarray(objs['data_x']), np.asarray(objs['data_D'])
        for j in range(k + 1):
            filled = np.flatnonzero(~np.isnan(dx[:, 0, j]))
            Ij = int(filled[0]) + 1 if filled.size else N
            h.rows.append((Ij, dx[Ij - 1:N, :, j].copy(), dD[Ij - 1:N, :, j].copy()))
        return h


def state_from_objs(objs):
    """the loop state after iteration k of a stored run, from the reference's dict (parareal.py:279-297)"""
    k = int(objs['k'])
    u, uG, uF = np.asarray(objs['u']), np.asarray(objs['uG']), np.asarray(objs['uF'])
    if u.ndim != 3:
        raise Exception('not an intermediate dump in the reference layout (u must be [N+1, n, k+2])')
    err = np.asarray(objs['err'])
    N = u.shape[0] - 1
    err_full = np.full((N + 1, N), np.nan)
    err_full[:, :err.shape[1]] = err
    err_full[:, k + 1:] = np.nan
    return dict(k=k, I=int(objs['I']), conv_int=[int(v) for v in objs.get('conv_int', [])], err=err_full,
                u=u[:, :, k + 1].copy(), uG=uG[:, :, k + 1].copy(), uF=uF[:, :, k].copy(),
                x=np.asarray(objs['x']).copy(), D=np.asarray(objs['D']).copy(),
                G_time=objs.get('G_time', 0), F_time=objs.get('F_time', 0))


class _Bag:
    """stand-in for a class of the reference that is not importable here: keeps the pickled attributes"""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Bag,), {'_foreign_module': module})


def load_dump(path):
    """unpickles an intermediate dump written by this package or by the reference (parareal.py:114-139)"""
    with open(path, 'rb') as fh:
        return _Unpickler(fh).load()


def adopt_model(mdl, n, N):
    """a model object of this package from whatever the dump holds (the reference's NNGP_p / BareParareal / GPjax_p,
    an attribute bag, or already one of ours): hyper-parameters, dataset, RNG state and timing accumulators are kept"""
    from .models import BareParareal, CudaNNGP, ModelAbstr
    from .gp_full import CudaGP
    if isinstance(mdl, ModelAbstr):
        return mdl
    name = getattr(mdl, 'name', None)
    a = mdl.__dict__
    if name == 'Parareal':
        new = BareParareal(N=N)
    elif name == 'NNGP':
        new = CudaNNGP(n=a.get('n', n), N=N, nn=a.get('nn', 'adaptive'), n_restarts=a.get('n_restarts', 1),
                       seed=a.get('seed', 45), fatol=a.get('fatol'), xatol=a.get('xatol'),
                       calc_detail_avg=a.get('calc_detail_avg', False), calc_parall_overhead=a.get('calc_parall_overhead', False))
        for key in ('rng', 'tot_train_t', 'train_count', 'detail_avg', 'overhead', 'k', 'x', 'y'):
            if key in a:
                setattr(new, key, a[key])
    elif name == 'GP':
        new = CudaGP(n=a.get('n', n), N=N, fatol=a.get('fatol'), xatol=a.get('xatol'))
        for key in ('rng', 'tot_train_t', 'train_count', 'k', 'x', 'y', 'hyp', 'thetas', 'jitters'):
            if key in a:
                setattr(new, key, a[key])
    else:
        raise Exception(f'cannot adopt a model named {name!r}')
    for key in ('train_time', 'pred_time', 'pred_times', 'time_k'):
        if key in a:
            setattr(new, key, a[key])
    return new
