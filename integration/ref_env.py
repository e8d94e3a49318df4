"""Imports the reference's own modules (systems, solver, RK, models, parareal, configs) UNMODIFIED.

Where the reference lives: $NNGP_REFERENCE_DIR, else /root/reference, else <repo>/baseline/_ref (a git-ignored copy
of the seven modules made by scripts/stage_reference.sh so that the binding can be exercised on a GPU box).

The reference imports jax, matplotlib, cycler and kiwisolver at module level (models.py:5-8, parareal.py:2,5,
solver.py:1, systems.py:2).  On a maintainer's machine those are installed and nothing here is used.  Where one of
them is missing (this image has none of them) a minimal stand-in is registered so that the modules import: the
binding overrides every method that would reach JAX (the fits and the propagators run in libnngpara.so), so the
stand-ins only have to survive `import` and class definition -- `jit` is the identity, `vmap` builds a closure that
raises if it is ever called.  This file is part of the binding, not of the test oracle."""
import importlib
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULES = ("utils", "systems", "configs", "RK", "solver", "models", "parareal")


def find_reference():
    for cand in (os.environ.get("NNGP_REFERENCE_DIR"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "parareal.py")) and os.path.isfile(os.path.join(cand, "models.py")):
            return cand
    return None


def _importable(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def _stand_ins():
    import numpy as np
    mods = {}
    if not _importable("jax"):
        jax = types.ModuleType("jax")
        jnp = types.ModuleType("jax.numpy")
        for name in dir(np):
            if not name.startswith("__"):
                setattr(jnp, name, getattr(np, name))

        def jit(f=None, **kw):
            return (lambda g: g) if f is None else f

        def vmap(f, *a, **k):
            def never(*args, **kwargs):
                raise RuntimeError("jax is not installed: this code path of the reference is replaced by libnngpara.so")
            return never

        class _Config:
            def update(self, *a, **k):
                pass

        jax.jit, jax.vmap, jax.numpy, jax.config = jit, vmap, jnp, _Config()
        jax.lax = types.ModuleType("jax.lax")
        jax.scipy = types.ModuleType("jax.scipy")
        jax.scipy.linalg = types.ModuleType("jax.scipy.linalg")
        cfg = types.ModuleType("jax.config")
        cfg.config = jax.config
        mods.update({"jax": jax, "jax.numpy": jnp, "jax.lax": jax.lax, "jax.scipy": jax.scipy,
                     "jax.scipy.linalg": jax.scipy.linalg, "jax.config": cfg})
    if not _importable("matplotlib"):
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        mods.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    if not _importable("cycler"):
        cyc = types.ModuleType("cycler")
        cyc.cycler = lambda *a, **k: None
        mods["cycler"] = cyc
    if not _importable("kiwisolver"):
        kiwi = types.ModuleType("kiwisolver")
        kiwi.Solver = object
        mods["kiwisolver"] = kiwi
    return mods


_ns = None


def import_reference(ref_dir=None):
    """namespace with the reference's modules; raises if the reference is not on this machine"""
    global _ns
    if _ns is not None:
        return _ns
    ref_dir = ref_dir or find_reference()
    if ref_dir is None:
        raise ImportError("the reference (parareal.py, models.py, ...) was not found: set NNGP_REFERENCE_DIR")
    for name, mod in _stand_ins().items():
        sys.modules.setdefault(name, mod)
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    ns = types.SimpleNamespace(ref_dir=ref_dir)
    for name in MODULES:
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(os.path.abspath(ref_dir)):
            del sys.modules[name]
        setattr(ns, name, importlib.import_module(name))
    _ns = ns
    return ns
