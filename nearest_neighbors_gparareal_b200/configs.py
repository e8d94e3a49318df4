"""Host mirror of reference configs.py:6-181 (`Config`): per-system presets with per-slice
Ng / Nf.  Read-only input of the hot path."""
import numpy as np

from .systems import (FHN_ODE, FHN_PDE, Rossler, Hopf, DblPend, Brusselator, Lorenz,
                      ThomasLabyrinth, Burgers, ODE)


class Config:
    def __init__(self, ode: ODE, N=None, d_x=None):
        if isinstance(ode, FHN_ODE):
            n = 40
            ng = n * 4
            config = dict(tspan=[0, 40], u0=np.array([-1, 1]), N=n, Ng=ng / n,
                          Nf=int(160000 / 160 * ng) / n, G='RK2', F='RK4')
        elif isinstance(ode, Rossler):
            n, ng, nf = 20 * 2, 45000 * 2, 2250000 * 2
            config = dict(tspan=[0, 170 * 2], u0=np.array([0, -6.78, 0.02]), N=n, Ng=ng / n, Nf=nf / n,
                          G='RK1', F='RK4')
        elif isinstance(ode, Hopf):
            if N is None:
                raise Exception('N must be provided')
            ng = 2 * 1024
            config = dict(tspan=[-20, 500], u0=np.array([0.1, 0.1, -20]), N=N, Ng=ng / N, Nf=ng * 85 / N,
                          G='RK1', F='RK8')
            ode.name += f'_{N}'
        elif isinstance(ode, DblPend):
            n = 32
            ng = 3072 + n
            config = dict(tspan=[0, 80], u0=np.array([-0.5, 0, 0, 0]), N=n, Ng=ng / n, Nf=ng * 70 / n,
                          G='RK1', F='RK8')
        elif isinstance(ode, Brusselator):
            n = 25
            ng = n * 10
            config = dict(tspan=[0, 100], u0=np.array([1, 3.07]), N=n, Ng=ng / n, Nf=ng * 100 / n,
                          G='RK4', F='RK4')
        elif isinstance(ode, Lorenz):
            n = 50
            ng = n * 6
            config = dict(tspan=[0, 18], u0=np.array([-15, -15, 20]), N=n, Ng=ng / n, Nf=ng * 75 / n,
                          G='RK4', F='RK4')
        elif isinstance(ode, ThomasLabyrinth):
            tot = {32: 10, 64: 10, 128: 40, 256: 100, 512: 100}
            if N not in tot:
                raise Exception('Invalid N value')
            ng = N * 10
            nf = ng * int(np.ceil(1e6 / ng))
            config = dict(tspan=[0, tot[N]], u0=np.array([4.6722764, 5.2437205e-10, -6.4444208e-10]), N=N,
                          Ng=ng / N, Nf=nf / N, G='RK1', F='RK4')
            ode.name += f'_{N}'
        elif isinstance(ode, FHN_PDE):
            config = self.fhn_pde(d_x)
        elif isinstance(ode, Burgers):
            # not in the reference's Config; values of Burgers.py:27-57 (light F of
            # Burgers_perf_across_m.py:30-31 when N is given as a small number of slices)
            n = 128 if N is None else N
            config = dict(tspan=[0, 5.9], N=n, Ng=4, Nf=40000, G='RK1', F='RK8')
        else:
            raise Exception('No config for input ODE')
        if 'u0' in config:
            ode.set_default_init_cond(config['u0'])
        self.config = config

    def fhn_pde(self, dx, *args, **kwargs):
        N = 512
        table = {10: (3, 150, 'RK2'), 12: (12, 550, 'RK2'), 14: (25, 950, 'RK2'), 16: (25, 1100, 'RK4')}
        mul, T, G = table.get(dx, (25, 1100, 'RK4'))
        Ng = N * mul
        Nf = int(np.ceil(1e4 / Ng) * Ng)
        return {'tspan': [0, T], 'N': N, 'Ng': Ng / N, 'Nf': Nf / N, 'G': G, 'F': 'RK8'}

    def _enforce_types(self, config):
        for key, val in config.items():
            if key in ['N', 'Ng', 'Nf']:
                config[key] = int(val)
            elif key in ['u0']:
                config[key] = np.array(val)
        return config

    def get(self):
        return self._enforce_types(self.config)
