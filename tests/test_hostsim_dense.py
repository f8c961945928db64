"""The dense-matrix form of the reference's optimiser (csrc/t2fit_lbfgsb_dense.cuh: L-BFGS-B with the limited-memory matrix
held as an n x n matrix, n <= 3) compiled for the host and held to the SAME parity thresholds as the compact form
(tests/test_hostsim_lbfgsb.py): the golden fixtures of the unmodified reference, judged against the reference's own
reproducibility floor (tests/golden/make_jitter.py)."""
import warnings

import numpy as np
import pytest
from scipy.optimize import minimize

from tests import hostsim
from tests.conftest import assert_lbfgsb_parity, fit_params_of, lbfgsb_parity_report, load_golden
from tests.test_hostsim_lbfgsb import _fg_floor, _fg_mono

FIXTURES = ["c1_gaussian_noprior", "c1_gaussian_prior", "c2_gaussian_noprior", "c2_gaussian_hf_prior", "c4_gaussian_noprior",
            "c3_floor_noprior", "c3_floor_prior", "c5_floor_noprior", "c3_rician_prior", "norm_gaussian",
            "cli3_gaussian_lf_noprior", "cli3_floor_hf_prior", "cli3_rician_hf_prior", "cli3_rician_lf_noprior"]


def _run(g, dense, **kw):
    fp = fit_params_of(g)
    return hostsim.lbfgsb(g["rows"], g["te"], g["fit"], g["x0"], g["bounds"], g["prior"], g["norm"], options=fp["options"],
                          dense=dense, **kw)


@pytest.mark.parametrize("name", FIXTURES)
def test_dense_form_reproduces_reference_fixtures(name):
    g = load_golden(name)
    o = _run(g, True, trace_cap=64)
    rep = lbfgsb_parity_report(o["x"][:, 1], o["nit"], o["status"] == 0, g)
    assert_lbfgsb_parity(rep, name + " (dense)")
    nt = g["trace_len"].shape[0]
    for i in range(nt):                                                  # callback traces, as for the compact form
        if o["nit"][i] != g["ref_nit"][i] or not g["reproducible"][i]:
            continue
        n = int(g["trace_len"][i])
        assert o["trace_len"][i] == n
        assert np.allclose(o["trace_f"][i, :min(n, 3)], g["trace_f"][i, :min(n, 3)], rtol=5e-3)
        assert np.allclose(o["trace_f"][i, n - 1], g["trace_f"][i, n - 1], rtol=1e-2)


@pytest.mark.parametrize("name", ["c3_floor_noprior", "c5_floor_noprior", "c3_floor_prior"])
def test_dense_form_restarts_where_the_compact_form_breaks_down(name, monkeypatch):
    """Every factorization breakdown of the compact form on the fixtures is the structural one (a pair stored while no
    variable was free never reaches the K matrix); the dense form refreshes its memory at the same event.  The two runs
    differ in rounding, so a handful of voxels may reach the event in one run only."""
    g = load_golden(name)
    monkeypatch.setenv("HOSTSIM_BRK", "1")        # the host build reports the first breakdown above the result code
    oc, od = _run(g, False), _run(g, True)
    bc, bd = ((oc["result"] >> 8) & 0xff), ((od["result"] >> 8) & 0xff)
    assert not np.any(bc & ~2 & 0x0f), "compact form: a breakdown that is not formk's"
    assert not np.any(bd & ~2 & 0x0f)
    bc, bd = (bc & 2) > 0, (bd & 2) > 0
    if name == "c3_floor_prior":
        assert bc.sum() == 0 and bd.sum() == 0
    else:
        assert bc.sum() >= 30
        assert (bc & bd).sum() >= 0.9 * max(bc.sum(), bd.sum())
    assert np.mean((oc["result"] & 0xff) == (od["result"] & 0xff)) >= 0.98      # same stopping test fired


@pytest.mark.parametrize("name,m", [("c2_gaussian_noprior", 150), ("c3_floor_prior", 100)])
def test_dense_core_follows_scipy_with_an_analytic_gradient(name, m):
    """No finite-difference noise: the dense form must walk scipy's path up to rounding in the quasi-Newton products
    (looser than the compact form's 1e-8, which repeats scipy's operations)."""
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
    o = hostsim.lbfgsb(rows, te, g["fit"], fp["initial_guess"], fp["param_bounds"], g["prior"], options=fp["options"], tol=-1.0,
                       dense=True)
    n = sx.shape[1]
    rel = np.abs(o["x"][:, :n] - sx).max(axis=1) / np.abs(sx).max(axis=1)
    print(name, "nit equal", np.mean(o["nit"] == snit), "within 1e-6", np.mean(rel <= 1e-6), "median", np.median(rel))
    assert np.mean(o["nit"] == snit) >= 0.95
    assert np.mean(rel <= 1e-6) >= 0.95
    assert np.median(rel) <= 1e-9


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
@pytest.mark.parametrize("prior", [True, False], ids=["prior", "noprior"])
def test_dense_edge_cases(fit, prior):
    g = load_golden(f"edge_{fit}_{'prior' if prior else 'noprior'}")
    o = _run(g, True)
    raises = np.array([len(str(e)) > 0 for e in g["ref_error"]])
    assert np.array_equal(o["status"] == 3, raises)                      # scipy's ValueError rows
    keep = ~raises
    assert np.array_equal((o["status"] == 0)[keep], g["ref_success"][keep])
    failed = keep & ~g["ref_success"]
    n = g["ref_params"].shape[1]
    assert np.allclose(o["x"][failed, :n], g["ref_params"][failed])      # clipped x0
    oc = _run(g, False)                                                  # and the compact form's answers on the rest
    ok = keep & g["ref_success"] & g["converged"] & (g["ref_params"][:, 1] > 10.0)
    rel = np.abs(o["x"][ok, 1] - oc["x"][ok, 1]) / oc["x"][ok, 1]
    assert np.mean(rel <= 1e-3) >= 0.75 and rel.max() <= 5e-2


def test_echo_quotient_from_one_reciprocal_is_the_ieee_division():
    """lb::EchoDiv: a / t2 for many numerators and one denominator as q0 = RN(a y), q = RN(q0 + (a - t2 q0) y) with
    y = RN(1 / t2) must be the correctly rounded quotient, bit for bit (the objective's exp(-2 TE / T2) argument), on the
    operand ranges the fit sees, on mantissas of all ones / a single one, and must fall back to the division itself near the
    ends of the exponent range.  (The 4e8-pair run quoted in DESIGN.md section 4 used the same recurrence in C.)"""
    import ctypes as C
    rng = np.random.default_rng(5)
    n = 400_000
    te = -2.0 * rng.integers(1, 4000, n) * 0.5
    t2 = np.exp(rng.uniform(np.log(1e-3), np.log(1e4), n))
    a = np.concatenate([te, -rng.uniform(10.0, 2000.0, n), -(1.0 + rng.random(n)), np.array([0.0, -0.0, -2.0, -1e-300, -1e300, -3.0])])
    b = np.concatenate([t2, rng.uniform(10.0, 2000.0, n), np.nextafter(2.0, 0.0) - rng.integers(0, 4, n) * 2.0 ** -52,
                        np.array([37.0, 37.0, 1e-310, 7.0, 1e-200, 1e250])])
    q = np.empty_like(a)
    L = hostsim.lib(strict=True)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    L.hostsim_echodiv(vp(a), vp(b), vp(q), C.c_int64(a.size), C.c_int(1))
    with np.errstate(all="ignore"):
        want = a / b
    assert np.array_equal(q, want), f"{np.count_nonzero(q != want)} of {a.size} quotients differ from a / b"
    # (array_equal: a zero numerator gives +0.0 where the division gives -0.0 -- TE = 0; exp of either is 1)


def test_constant_bank_exponential_is_within_one_ulp():
    """lb::exp_echo (the T2_DENSE_EXP experiment: Taylor degree 13 after the two-part ln 2 reduction; off by default, measured
    slower than the library's exp on the GPU) against an 80-bit reference on the arguments the echo loop produces."""
    import ctypes as C
    rng = np.random.default_rng(1)
    a = np.concatenate([-rng.uniform(0, 700, 200000), -np.exp(rng.uniform(np.log(1e-12), np.log(700), 200000)),
                        np.array([0.0, -0.0, -699.9, -700.0, -745.0, -1e9, 3.0, -np.inf])])
    out = np.empty_like(a)
    hostsim.lib(strict=True).hostsim_exp_echo(a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_int64(a.size))
    fin = np.isfinite(a) & (a > -700)
    ref = np.exp(a[fin].astype(np.longdouble))
    err = np.abs((out[fin].astype(np.longdouble) - ref) / np.spacing(np.exp(a[fin])))
    assert float(err.max()) <= 1.0
    assert np.array_equal(out[~fin], np.exp(a[~fin]))                    # outside (-700, 700): the library's own
