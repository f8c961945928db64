"""The reference-faithful solver (csrc/t2fit_lbfgsb.cuh) compiled for the host, checked on GPU-less CI:
(1) its optimiser core against scipy's L-BFGS-B itself, driven by an analytic gradient (no finite-difference
noise, so trajectories must agree to rounding); (2) the full emulation -- scipy's 2-point differences, the
reference's objectives in numpy operation order -- against the golden fixtures of the unmodified reference,
judged against the reference's own reproducibility floor (tests/golden/make_jitter.py)."""
import warnings

import numpy as np
import pytest
from scipy.optimize import minimize
from scipy.special import i0e

from tests import hostsim
from tests.conftest import assert_lbfgsb_parity, fit_params_of, lbfgsb_parity_report, load_golden


def _fg_mono(p, te, y):
    u = np.exp(-te / p[1]); m = p[0] * u; r = y - m
    return np.sum(r * r) / len(te), np.array([np.sum(-2 * r * u), np.sum(-2 * r * m * te / p[1] ** 2)]) / len(te)


def _fg_floor(p, te, y):
    u2 = np.exp(-2 * te / p[1]); m = np.sqrt(p[0] ** 2 * u2 + p[2] ** 2); r = y - m
    return np.sum(r * r) / len(te), np.array([np.sum(-2 * r * p[0] * u2 / m), np.sum(-2 * r * p[0] ** 2 * u2 * te / p[1] ** 2 / m),
                                              np.sum(-2 * r * p[2] / m)]) / len(te)


@pytest.mark.parametrize("name,m", [("c2_gaussian_noprior", 150), ("c1_gaussian_prior", 100), ("c3_floor_noprior", 150),
                                   ("c3_floor_prior", 100)])
def test_optimiser_core_follows_scipy_lbfgsb(name, m):
    """Same f and g on both sides (analytic gradient) -> the restated L-BFGS-B (Cauchy point, subspace step,
    More'-Thuente search, compact updates, memory refresh on a failed Cholesky) must walk scipy's path."""
    g = load_golden(name)
    fp = fit_params_of(g)
    fg = _fg_mono if g["fit"] == "gaussian" else _fg_floor
    rows, te = g["rows"][:m], g["te"].astype(float)
    sx, snit = [], []
    for i in range(m):
        bounds = list(fp["param_bounds"])
        if not g["prior"]:
            bounds[0] = (float(rows[i, 0]), 10000.0); bounds[1] = (10.0, 2000.0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = minimize(fg, fp["initial_guess"], args=(te, rows[i].astype(float)), method="L-BFGS-B", bounds=bounds,
                         options={k: v for k, v in fp["options"].items() if k != "disp"}, jac=True)
        sx.append(r.x); snit.append(r.nit)
    sx, snit = np.array(sx), np.array(snit)
    o = hostsim.lbfgsb(rows, te, g["fit"], fp["initial_guess"], fp["param_bounds"], g["prior"], options=fp["options"], tol=-1.0)
    n = sx.shape[1]
    rel = np.abs(o["x"][:, :n] - sx).max(axis=1) / np.abs(sx).max(axis=1)
    assert np.mean(o["nit"] == snit) >= 0.97, (name, np.mean(o["nit"] == snit))
    assert np.mean(rel <= 1e-8) >= 0.97, (name, np.mean(rel <= 1e-8))
    assert np.median(rel) <= 1e-12


@pytest.mark.parametrize("name", ["c1_gaussian_noprior", "c1_gaussian_prior", "c2_gaussian_noprior", "c2_gaussian_hf_prior", "c4_gaussian_noprior",
                                  "c3_floor_noprior", "c3_floor_prior", "c5_floor_noprior", "c3_rician_prior", "norm_gaussian",
                                  "cli3_gaussian_lf_noprior", "cli3_floor_hf_prior", "cli3_rician_hf_prior", "cli3_rician_lf_noprior"])
def test_emulation_reproduces_reference_fixtures(name):
    g = load_golden(name)
    fp = fit_params_of(g)
    o = hostsim.lbfgsb(g["rows"], g["te"], g["fit"], g["x0"], g["bounds"], g["prior"], g["norm"], options=fp["options"],
                       trace_cap=64)
    rep = lbfgsb_parity_report(o["x"][:, 1], o["nit"], o["status"] == 0, g)
    assert_lbfgsb_parity(rep, name)
    # callback traces of the first voxels (run_t2mapping.py:180-234)
    nt = g["trace_len"].shape[0]
    for i in range(nt):
        if o["nit"][i] != g["ref_nit"][i] or not g["reproducible"][i]:
            continue
        n = int(g["trace_len"][i])
        assert o["trace_len"][i] == n
        # the first iterations agree closely, the last one (the answer) too; in between finite-difference noise lets
        # the two trajectories drift by a few percent before they contract to the same point
        assert np.allclose(o["trace_f"][i, :min(n, 3)], g["trace_f"][i, :min(n, 3)], rtol=5e-3)
        assert np.allclose(o["trace_f"][i, n - 1], g["trace_f"][i, n - 1], rtol=1e-2)
        assert np.isnan(o["trace_step"][i, 0])


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
@pytest.mark.parametrize("prior", [True, False], ids=["prior", "noprior"])
def test_emulation_edge_cases(fit, prior):
    g = load_golden(f"edge_{fit}_{'prior' if prior else 'noprior'}")
    fp = fit_params_of(g)
    o = hostsim.lbfgsb(g["rows"], g["te"], fit, g["x0"], g["bounds"], prior, options=fp["options"])
    raises = np.array([len(str(e)) > 0 for e in g["ref_error"]])
    assert np.array_equal(o["status"] == 3, raises)                      # scipy's ValueError rows
    keep = ~raises
    assert np.array_equal((o["status"] == 0)[keep], g["ref_success"][keep])
    failed = keep & ~g["ref_success"]
    n = g["ref_params"].shape[1]
    assert np.allclose(o["x"][failed, :n], g["ref_params"][failed])      # clipped x0
    ok = keep & g["ref_success"] & g["converged"] & (g["ref_params"][:, 1] > 10.0)
    rel = np.abs(o["x"][ok, 1] - g["ref_params"][ok, 1]) / g["ref_params"][ok, 1]
    # pathological rows: a rounding-level change moves the reference's own stopping point by percents
    assert np.mean(rel <= 1e-3) >= 0.75 and rel.max() <= 5e-2


def test_scaled_bessel_i0e():
    xs = np.concatenate([np.linspace(0, 8, 2001), np.geomspace(8, 1e9, 2001), -np.linspace(0, 40, 50)])
    mine = np.array([hostsim.lib().hostsim_i0e(float(x)) for x in xs])
    ref = i0e(xs)
    assert np.max(np.abs(mine - ref) / ref) < 1e-15
