"""Host mirror of the reference's `ODE` protocol and system catalogue (systems.py:23-459).

Same class names, constructor arguments, attributes (`name`, `normalizer`, `u0`) and methods
(`get_vector_field`, `get_init_cond`, `set_default_init_cond`, `get_dim`).  The arithmetic
of every vector field lives in csrc/rk.cu; a Python object only carries the system id, its
parameters and the normalisation bounds to the device (`device_desc`).  `get_vector_field()`
returns a callable f(t, u) that evaluates the field ON THE GPU (nngp_rhs_eval_host) -- there is
no NumPy twin of the right-hand sides in this package.
"""
import numpy as np

from . import _lib
from .utils import Normalize


class DeviceVectorField:
    """f(t, u) of systems.py:32-44; also what CudaSolverRK takes in place of the closure."""

    def __init__(self, ode):
        self.ode = ode

    def __call__(self, t, u):
        u = np.asarray(u, dtype=float)
        h, sys = self.ode.device_system()
        return h.rhs_eval_host(sys, u)


class ODE():
    system_key = None  # key into _lib.SYSTEM_IDS

    def __init__(self, name, mn, mx, u0, normalization=None, use_jax=True):
        self.name = name
        self.normalizer = Normalize(np.asarray(mn, dtype=float), np.asarray(mx, dtype=float), normalization)
        self.u0 = self.normalizer.fit(np.asarray(u0, dtype=float))
        self.use_jax = use_jax  # accepted for signature compatibility; there is one (device) path
        self._dev = {}

    # -- reference protocol --------------------------------------------------------------
    def get_vector_field(self):
        return DeviceVectorField(self)

    def _get_f(self, use_jax):
        raise NotImplementedError('the un-normalised host field does not exist in the device package')

    def set_default_init_cond(self, u0):
        self.u0 = self.normalizer.fit(np.asarray(u0, dtype=float))

    def get_init_cond(self, *args, u0=None, **kwargs):
        if u0 is None:
            u0 = self.u0
        else:
            u0 = self.normalizer.fit(np.asarray(u0, dtype=float))
        return np.array(u0, dtype=float)

    def get_dim(self):
        return self.u0.shape[0]

    # -- device side ---------------------------------------------------------------------
    def device_params(self):
        """parameters of the field, computed here exactly as the reference computes them"""
        return np.zeros(0)

    def device_desc(self):
        if self.system_key is None:
            raise NotImplementedError('This is an abstract class')
        normalize = self.normalizer.norm_type == '-11'
        return dict(system_id=_lib.SYSTEM_IDS[self.system_key], d=self.get_dim(),
                    params=self.device_params(), normalize=normalize,
                    mn=self.normalizer.mn, mx=self.normalizer.mx)

    def device_system(self, handle=None):
        """(handle, system index) -- created once per handle (nngp_system_create)"""
        h = handle or _lib.default_handle()
        key = id(h)
        if key not in self._dev:
            d = self.device_desc()
            self._dev[key] = h.system_create(d['system_id'], d['d'], d['params'], d['normalize'], d['mn'], d['mx'])
        return h, self._dev[key]

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_dev'] = {}
        return state


class FHN_ODE(ODE):
    system_key = 'FHN_ODE'

    def __init__(self, **kwargs):
        mn, mx = np.array([[-2, -1], [2.1, 1.2]])
        super().__init__('FHN_ODE', mn, mx, np.array([-1, 1]), **kwargs)


class Rossler(ODE):
    system_key = 'Rossler'

    def __init__(self, **kwargs):
        mn, mx = np.array([[-10, -11, 0], [12, 8, 23]])
        super().__init__('Rossler', mn, mx, np.array([0, -6.78, 0.02]), **kwargs)


class Hopf(ODE):
    system_key = 'Hopf'

    def __init__(self, tspan=[-20, 500], **kwargs):
        mn, mx = np.array([[-23, -23, 0], [23, 23, 1]])
        self.maxtime = tspan[1]
        super().__init__('Hopf', mn, mx, np.array([0.1, 0.1, tspan[0]]), **kwargs)

    def device_params(self):
        return np.array([float(self.maxtime)])


class DblPend(ODE):
    system_key = 'DblPend'

    def __init__(self, **kwargs):
        mn, mx = np.array([[-2, -2.5, -17, -3.5], [2, 2.5, 1, 3.5]])
        super().__init__('DblPend', mn, mx, np.array([-0.5, 0, 0, 0]), **kwargs)


class Brusselator(ODE):
    system_key = 'Brusselator'

    def __init__(self, **kwargs):
        mn, mx = np.array([[0.4, 0.9], [4, 5]])
        super().__init__('Brusselator', mn, mx, np.array([1, 3.07]), **kwargs)


class Lorenz(ODE):
    system_key = 'Lorenz'

    def __init__(self, **kwargs):
        mn, mx = np.array([[-17.1, -23, 6], [18.1, 25, 45]])
        super().__init__('Lorenz', mn, mx, np.array([-15, -15, 20]), **kwargs)


class ThomasLabyrinth(ODE):
    system_key = 'ThomasLabyrinth'

    def __init__(self, **kwargs):
        mn, mx = np.array([[-12, -12, -12], [12, 12, 12]])
        u0 = np.array([4.6722764, 5.2437205e-10, -6.4444208e-10])
        super().__init__('ThomasLabyrinth', mn, mx, u0, **kwargs)


class FHN_PDE(ODE):
    """FitzHugh-Nagumo on a periodic d_x x d_x grid (systems.py:291-398)."""
    system_key = 'FHN_PDE'

    def __init__(self, d_x, seed=45, **kwargs):
        self.d_x = d_x
        self.d_y = d_x
        d = 2 * (d_x * d_x)
        self.d = d
        mn, mx = np.array([[-1] * d, [1] * d])
        # same legacy global stream as the reference (systems.py:303-312)
        np.random.seed(seed)
        rng = np.random.Generator(np.random.get_bit_generator())
        u0 = rng.uniform(size=self.d)
        super().__init__(f'FHN_PDE_{d_x}', mn, mx, u0, **kwargs)

    def device_params(self):
        # entries of a*(DXX+DYY) and b*(DXX+DYY) as systems.py:321-353,365-366 produce them
        d_x = self.d_x
        dx = (1 - (-1)) / (d_x - 1)
        off = (1 / (dx ** 2)) * 1.0
        diag = (1 / (dx ** 2)) * (-2.0)
        lap_diag = diag + diag       # DXX[p,p] + DYY[p,p]
        lap_off = off + 0.0          # one of the two Kronecker terms is zero off the diagonal
        a, b, k, tau = 2.8E-4, 5E-3, -5E-3, 0.1
        return np.array([float(d_x), a * lap_diag, a * lap_off, b * lap_diag, b * lap_off, k, 1 / tau])


class Burgers(ODE):
    """Viscous Burgers, periodic central differences (systems.py:402-459)."""
    system_key = 'Burgers'

    def __init__(self, d_x, nu=1 / 100, **kwargs):
        self.d_x = d_x
        self.nu = nu
        d = d_x
        self.d = d
        mn, mx = np.array([[0] * d, [1] * d])
        x_fine = np.linspace(-1, 1, num=(d - 1) + 1)
        u0 = 0.5 * (np.cos(4.5 * np.pi * x_fine) + 1)
        super().__init__(f'Burgers_{d_x}', mn, mx, u0, **kwargs)

    def device_params(self):
        d, nu = self.d, self.nu
        dx = (1 - (-1)) / (d - 1)
        dxx_off = (nu / (dx ** 2)) * 1.0
        dxx_diag = (nu / (dx ** 2)) * (-2.0)
        dx_off = (1 / (2 * dx)) * 1.0
        return np.array([dxx_off, dxx_diag, dx_off])
