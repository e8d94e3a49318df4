"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Container-only helper: imports the UNMODIFIED reference modules from
/root/reference (read-only, absent on the GPU box) under stub modules for the
packages the reference imports but this image lacks (jax, matplotlib, cycler,
kiwisolver).  Used by oracle/make_golden.py to pin the oracle restatement and
to generate the fixtures in tests/golden/.  Nothing in tests -m gpu, smoke()
or bench.py may import this file.

The `jax` stand-in is NumPy backed: jit = identity, vmap = python loops over
the mapped axis, jnp = numpy (with linalg.cholesky returning NaNs instead of
raising, which is what XLA's Cholesky does), jax.scipy.linalg.solve_triangular
= SciPy's.  With use_jax=False the reference takes its own `_f_np` /
`_RK_numpy_` branches (systems.py:47-51, RK.py:97-98,107-108), so the only
reference code running through the stand-in is models.py (kernel, Cholesky,
triangular solves) -- i.e. the "reference's own NumPy/SciPy path".
"""
import os
import sys
import types

import numpy as np
import scipy.linalg

REF_DIR = os.environ.get("NNGP_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "models.py"))


def _make_jax():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if not name.startswith("__"):
            setattr(jnp, name, getattr(np, name))
    jlinalg = types.ModuleType("jax.numpy.linalg")
    for name in dir(np.linalg):
        if not name.startswith("__"):
            setattr(jlinalg, name, getattr(np.linalg, name))

    def cholesky(a):
        try:
            return np.linalg.cholesky(a)
        except np.linalg.LinAlgError:
            out = np.empty_like(np.asarray(a, dtype=float))
            out.fill(np.nan)
            return out

    jlinalg.cholesky = cholesky
    jnp.linalg = jlinalg

    def jit(f=None, static_argnums=None, **kw):
        if f is None:
            return lambda g: g
        return f

    def vmap(f, in_axes=0, out_axes=0):
        def mapped(*args):
            axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
            n = None
            for a, ax in zip(args, axes):
                if ax is not None:
                    n = np.shape(a)[ax]
                    break
            outs = []
            for i in range(n):
                call = [np.take(a, i, axis=ax) if ax is not None else a for a, ax in zip(args, axes)]
                outs.append(f(*call))
            return np.stack(outs, axis=out_axes)
        return mapped

    class _Config:
        def update(self, *a, **k):
            pass

    lax = types.ModuleType("jax.lax")

    def fori_loop(lo, hi, body, init):
        val = init
        for i in range(int(lo), int(hi)):
            val = body(i, val)
        return val

    lax.fori_loop = fori_loop
    jscipy = types.ModuleType("jax.scipy")
    jsl = types.ModuleType("jax.scipy.linalg")

    def solve_triangular(a, b, lower=False, **kw):
        if np.any(np.isnan(a)):
            out = np.empty_like(np.asarray(b, dtype=float))
            out.fill(np.nan)
            return out
        return scipy.linalg.solve_triangular(a, b, lower=lower, check_finite=False)

    jsl.solve_triangular = solve_triangular
    jscipy.linalg = jsl
    jax.jit = jit
    jax.vmap = vmap
    jax.numpy = jnp
    jax.lax = lax
    jax.scipy = jscipy
    jax.config = _Config()
    cfgmod = types.ModuleType("jax.config")
    cfgmod.config = jax.config
    return {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jlinalg, "jax.lax": lax,
            "jax.scipy": jscipy, "jax.scipy.linalg": jsl, "jax.config": cfgmod}


def _make_plot_stubs():
    mods = {}
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")

    def _noop(*a, **k):
        raise RuntimeError("matplotlib is not installed; plotting is out of scope")

    plt.subplots = _noop
    plt.figure = _noop
    mpl.pyplot = plt
    mods["matplotlib"] = mpl
    mods["matplotlib.pyplot"] = plt
    cyc = types.ModuleType("cycler")
    cyc.cycler = lambda *a, **k: None
    mods["cycler"] = cyc
    kiwi = types.ModuleType("kiwisolver")
    kiwi.Solver = object
    mods["kiwisolver"] = kiwi
    return mods


_loaded = None


def load_reference(fast_kernel=True):
    """Returns a namespace with the reference's modules (systems, configs, solver,
    RK, models, parareal).  fast_kernel=True replaces NNGP_p.kernel_jit (a nested
    vmap of k_gauss, models.py:145-155) by the algebraically identical vectorised
    expression -- the python-loop vmap costs ~10x; SURVEY.md section 8c records that both
    give the same K/conv_int.  fast_kernel=False keeps the literal nested vmap."""
    global _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_DIR} (it only exists in the build container)")
    if _loaded is None:
        for name, mod in {**_make_jax(), **_make_plot_stubs()}.items():
            sys.modules.setdefault(name, mod)
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import importlib
        ns = types.SimpleNamespace()
        for name in ("utils", "systems", "configs", "RK", "solver", "models", "parareal"):
            if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(REF_DIR):
                del sys.modules[name]
            setattr(ns, name, importlib.import_module(name))
        ns._literal_kernel = ns.models.NNGP_p.kernel_jit
        _loaded = ns
    ns = _loaded
    if fast_kernel:
        from scipy.spatial.distance import cdist

        def kernel_vec(x, y, kernel_params):
            sigma_x, sigma_y = kernel_params
            y2 = np.atleast_2d(y)
            r2 = cdist(np.atleast_2d(x), y2, metric="sqeuclidean")
            return 10 ** (sigma_y) * np.exp(-0.5 * (1 / (10 ** sigma_x)) * r2)

        ns.models.NNGP_p.kernel_jit = staticmethod(kernel_vec)
    else:
        ns.models.NNGP_p.kernel_jit = ns._literal_kernel
    return ns
