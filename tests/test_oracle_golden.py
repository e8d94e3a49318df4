"""CPU: pins the oracle (oracle/*.py) against outputs of the UNMODIFIED reference recorded in
tests/golden/ by oracle/make_golden.py, and against the installed SciPy where the reference
calls SciPy.  No GPU, no /root/reference needed."""
import json
import os

import numpy as np
import pytest
from scipy.optimize import minimize
from scipy.spatial.distance import cdist

from oracle import nelder_mead as onm
from oracle import nngp as onn
from oracle import parareal as opara
from oracle import rk as ork
from oracle import systems as osys
from helpers import load_run, samples, case_system, oracle_system, GOLDEN

RK_SYSTEMS = {
    "lorenz": lambda: osys.Lorenz(normalization="-11"), "lorenz_id": lambda: osys.Lorenz(),
    "hopf": lambda: osys.Hopf(normalization="-11"), "rossler": lambda: osys.Rossler(normalization="-11"),
    "fhn_ode": lambda: osys.FHN_ODE(normalization="-11"), "brusselator": lambda: osys.Brusselator(normalization="-11"),
    "dblpend": lambda: osys.DblPend(normalization="-11"), "thomas": lambda: osys.ThomasLabyrinth(normalization="-11"),
    "burgers128": lambda: osys.Burgers(d_x=128, normalization="-11"), "burgers32": lambda: osys.Burgers(d_x=32, normalization="-11"),
    "fhn16": lambda: osys.FHN_PDE(d_x=16), "fhn4": lambda: osys.FHN_PDE(d_x=4),
    "fhn4_n": lambda: osys.FHN_PDE(d_x=4, normalization="-11"),
}


@pytest.fixture(scope="module")
def rkv():
    return np.load(os.path.join(GOLDEN, "rk_vectors.npz"))


@pytest.mark.parametrize("name", sorted(RK_SYSTEMS))
def test_vector_field_and_rk_bitwise(rkv, name):
    """oracle/systems.py + oracle/rk.py == reference systems.py `_f_np` + RK.py `_RK_numpy_`, bit for bit"""
    s = RK_SYSTEMS[name]()
    U = rkv[f"{name}_u"]
    assert np.array_equal(s.u0, U[0])
    got = np.stack([s.f(0.3, u) for u in U])
    assert np.array_equal(got, rkv[f"{name}_f"])
    for method in ("RK1", "RK2", "RK4", "RK8"):
        t0, t1, steps = rkv[f"{name}_{method}_t"]
        got = np.stack([ork.rk_last(s.f, method, t0, t1, int(steps), u) for u in U[:3]])
        assert np.array_equal(got, rkv[f"{name}_{method}_u1"]), method


FULL_SYSTEMS = {"lorenz": lambda: osys.Lorenz(normalization="-11"), "hopf": lambda: osys.Hopf(normalization="-11"),
                "burgers32": lambda: osys.Burgers(d_x=32, normalization="-11"), "fhn4": lambda: osys.FHN_PDE(d_x=4),
                "fhn6": lambda: osys.FHN_PDE(d_x=6)}


@pytest.mark.parametrize("name", sorted(FULL_SYSTEMS))
def test_rk_full_trajectory_bitwise(name):
    """oracle rk_full == reference RK.run (`_RK_numpy_`, every step; solver.py:109-113 run_F_full), bit for bit"""
    z = np.load(os.path.join(GOLDEN, "rk_full_vectors.npz"))
    s = FULL_SYSTEMS[name]()
    for method in ("RK1", "RK4", "RK8"):
        t0, t1, steps = z[f"{name}_{method}_t"]
        got = ork.rk_full(s.f, method, t0, t1, int(steps), z[f"{name}_u0"])
        want = z[f"{name}_{method}_traj"]
        assert got.shape == want.shape == (int(steps) + 1, z[f"{name}_u0"].shape[0])
        assert np.array_equal(got, want), method
        assert np.array_equal(got[-1], ork.rk_last(s.f, method, t0, t1, int(steps), z[f"{name}_u0"]))


def test_presets_match_reference_configs():
    """configs.py presets recorded with the golden runs"""
    z, cfg, _ = load_run("lorenz_N50_m11")
    p = osys.preset(osys.Lorenz(normalization="-11"))
    assert {k: p[k] for k in ("N", "Ng", "Nf", "G", "F", "tspan")} == {k: cfg[k] for k in ("N", "Ng", "Nf", "G", "F", "tspan")}
    z, cfg, _ = load_run("hopf_N32_m15")
    p = osys.preset(osys.Hopf(normalization="-11"), N=32)
    assert {k: p[k] for k in ("N", "Ng", "Nf", "G", "F", "tspan")} == {k: cfg[k] for k in ("N", "Ng", "Nf", "G", "F", "tspan")}
    p = osys.preset(osys.FHN_PDE(d_x=16))
    assert (p["N"], p["Ng"], p["Nf"], p["G"], p["F"], p["tspan"]) == (512, 25, 25, "RK4", "RK8", [0, 1100])


def test_sqdist_is_cdist_arithmetic():
    rng = np.random.default_rng(3)
    for d in (3, 128, 512):
        x = rng.standard_normal((200, d))
        q = rng.standard_normal(d)
        assert np.array_equal(onn.sqdist_rows(q, x), cdist(q[None], x, "sqeuclidean")[0])
    # default argsort == stable argsort on data without duplicate rows (tie rule)
    dist = onn.sqdist_rows(q, x)
    assert np.array_equal(np.argsort(dist), np.argsort(dist, kind="stable"))


def test_nelder_mead_restatement_equals_scipy():
    """oracle/nelder_mead.py == scipy.optimize.minimize(method='Nelder-Mead') on the reference's objective"""
    rng = np.random.default_rng(1)
    m, d = 11, 3
    x = rng.uniform(-1, 1, (200, d))
    y = 1e-3 * np.sin(x @ (rng.standard_normal((d, d)) / np.sqrt(d)))
    idx, _ = onn.knn(x[5] + 1e-3, x, m)
    r2 = onn.pairwise_sqdist(x[idx], x[idx])
    for j in range(d):
        for jit in onn.JITTERS[::2]:
            st = rng.integers(-8, 0, 2)
            f = lambda th: onn.neg_log_lik(r2, y[idx, j], th, jit)
            res = minimize(f, st, method="Nelder-Mead", options={"fatol": 0.1, "xatol": 0.1})
            xo, fo, ne, it, status = onm.nelder_mead(f, st, xatol=0.1, fatol=0.1)
            assert np.array_equal(xo, res.x) and fo == res.fun and ne == res.nfev and it == res.nit and status == res.status
    # budget exhaustion path (maxfev = 400): a noisy objective never converges
    noise = np.random.default_rng(0)
    g = lambda th: float(noise.standard_normal())
    noise2 = np.random.default_rng(0)
    g2 = lambda th: float(noise2.standard_normal())
    res = minimize(g, [-3.0, -4.0], method="Nelder-Mead", options={"fatol": 0.1, "xatol": 0.1})
    xo, fo, ne, it, status = onm.nelder_mead(g2, [-3.0, -4.0], xatol=0.1, fatol=0.1)
    assert np.array_equal(xo, res.x) and fo == res.fun and ne == res.nfev == 400 and status == res.status == 1


def test_select_rule():
    assert onn.select([3.0, 1.0, 1.0, 2.0]) == 1
    assert onn.select([-3.0, -10.0, -9.5, -10.0]) == 1
    assert onn.select([np.inf, np.inf]) == 0
    assert onn.select([5.0, np.inf, 4.0]) == 2
    assert (10.0 ** np.float64(-17.0)) == 1e-17 and all(10 ** j == l for j, l in zip(onn.JITTERS, [1e-20, 1e-19, 1e-18, 1e-17, 1e-16, 1e-15, 1e-14, 1e-13, 1e-12]))


@pytest.mark.parametrize("name", ["lorenz_N50_m11", "lorenz_N32_m11", "lorenz_N50_adaptive", "hopf_N32_m15",
                                  "burgers_d32_N32_m12", "fhn_d32_N32_m12"])
def test_predict_equals_reference_samples(name):
    """oracle/nngp.predict on the recorded (query, dataset prefix, starts) == NNGP_p.predict output.
    Exact in the generating container; ulp-level BLAS differences across CPUs are tolerated."""
    z, cfg, mkw = load_run(name)
    x, D = z["x"], z["D"]
    checked = 0
    pde = x.shape[1] > 8   # d = 32: 288 searches per predict, ~2 s each in the Python oracle
    for s in samples(z)[:3 if pde else 6]:
        n = int(s["n_rows"])
        got = onn.predict(s["query"], x[:n], D[:n], int(s["m"]), s["starts"].astype(np.int64))
        np.testing.assert_allclose(got, s["preds"], rtol=1e-6, atol=1e-13)
        checked += 1
    assert checked >= 3


def test_published_known_answers_recorded():
    with open(os.path.join(GOLDEN, "published.json")) as fh:
        pub = json.load(fh)
    fhn = pub["FHN_scal_times/FHN_scal_times_16_512_nngp"]["NNGP"]
    assert fhn["K"] == 6 and fhn["conv_int"] == [1, 2, 3, 4, 7, 512]
    rows = [r for r in pub["NNGP_all_but_pend"] if r[0] == "lorenz_n" and r[3] == "11" and r[2] == 5e-07]
    assert rows and all(r[1] == 10 for r in rows)
    # the shim-imported reference reproduced the published Lorenz and Hopf K
    assert int(load_run("lorenz_N50_m11")[0]["K"]) == 10
    assert int(load_run("hopf_N32_m15")[0]["K"]) == 9


def test_oracle_driver_reproduces_reference_run():
    """oracle/parareal.py (first 3 iterations of Lorenz N=32, m=11, seed 45) == the reference run"""
    z, cfg, mkw = load_run("lorenz_N32_m11")
    s = oracle_system("lorenz")
    solver = opara.OracleSolver(s.f, cfg["Ng"], cfg["Nf"], cfg["F"], cfg["G"])
    model = onn.OracleNNGP(n=3, N=cfg["N"], **mkw)
    out = opara.parareal(s.u0, solver, cfg["tspan"], cfg["N"], model, epsilon=float(z["epsilon"]), early_stop=3)
    assert out["conv_int"] == list(z["conv_int"][:3])
    np.testing.assert_allclose(out["err"][:, :3], z["err"][:, :3], rtol=1e-6, atol=1e-12, equal_nan=True)
    n = out["x"].shape[0]
    np.testing.assert_allclose(out["x"], z["x"][:n], rtol=1e-9, atol=1e-12)
