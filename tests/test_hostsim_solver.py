"""Solver LOGIC on the CPU: t2fit_core.cuh compiled with g++ (tests/hostsim) against the golden fixtures.

This is not the product path (the product is the sm_100a build, tested under -m gpu); it lets the
GPU-less CI catch regressions in bracketing / active-set / status logic.  libm exp2f stands in for
MUFU.EX2, so values differ from the device in the last bits only.
"""
import numpy as np
import pytest

from tests.conftest import load_golden
from tests import hostsim

T2_RTOL = 1e-3


def run(g, **kw):
    return hostsim.fit(g["rows"], g["te"], g["fit"], g["x0"], g["bounds"], g["prior"], g["norm"], **kw)


@pytest.mark.parametrize("use_double", [False, True], ids=["f32", "f64"])
@pytest.mark.parametrize("name", ["c1_gaussian_noprior", "c1_gaussian_prior", "c2_gaussian_noprior", "c2_gaussian_hf_prior"])
def test_mono2_reaches_bounded_minimum(name, use_double):
    g = load_golden(name)
    o = run(g, use_double=use_double)
    assert (o["status"] == 0).all()
    rel_e = np.abs(o["t2"] - g["exact_params"][:, 1]) / g["exact_params"][:, 1]
    assert rel_e.max() <= 1e-4
    rel = np.abs(o["t2"] - g["ref_params"][:, 1]) / g["ref_params"][:, 1]
    assert rel[g["converged"]].max() <= T2_RTOL
    assert o["nit"].max() <= 16 and o["nit"].mean() < 6


def test_preset_start_reaches_same_minimum():
    g = load_golden("c2_gaussian_noprior")
    a, b = run(g), run(g, init=1)
    assert (np.abs(a["t2"] - b["t2"]) / a["t2"]).max() < 1e-4


@pytest.mark.parametrize("name", ["c3_floor_noprior", "c3_floor_prior", "c5_floor_noprior"])
def test_floor3_not_worse_than_reference_point(name):
    g = load_golden(name)
    o = run(g)
    assert (o["status"] != 0).mean() < 0.01
    te, y = g["te"][None, :], g["rows"].astype(np.float64)

    def mse(k, t2, s):
        return ((y - np.sqrt(k[:, None] ** 2 * np.exp(-2 * te / t2[:, None]) + s[:, None] ** 2)) ** 2).mean(1)
    f_m = mse(o["k"].astype(float), o["t2"].astype(float), o["sigma"].astype(float))
    f_r = mse(*g["ref_params"].T)
    assert (f_m <= f_r * (1 + 1e-4) + 1e-6).mean() >= 0.85


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
@pytest.mark.parametrize("prior", [True, False])
def test_status_codes_match_reference_failed_sets(fit, prior):
    g = load_golden(f"edge_{fit}_{'prior' if prior else 'noprior'}")
    o = run(g)
    raised = np.array([len(str(e)) > 0 for e in g["ref_error"]])
    assert np.array_equal(o["status"] == 3, raised)                        # scipy ValueError <-> BADBOUNDS
    assert np.array_equal(o["status"][~raised] == 0, g["ref_success"][~raised])
    failed = (~g["ref_success"]) & ~raised
    np.testing.assert_allclose(o["t2"][failed], g["ref_params"][failed, 1], rtol=1e-6)
    np.testing.assert_allclose(o["k"][failed], g["ref_params"][failed, 0], rtol=1e-6)
    assert np.isnan(o["fun"][failed]).all() and (o["nit"][failed] == 0).all()


def test_bounds_validation_raises_like_scipy():
    g = load_golden("c1_gaussian_prior")
    with pytest.raises(ValueError, match="upper bound"):
        hostsim.fit(g["rows"][:4], g["te"], "gaussian", g["x0"], [(600, 100), (10, 600)], True)


def test_noise_free_signals_converge_in_one_or_two_passes():
    """At a converged point the Newton step rounds to zero (r + dr == r with fused multiply-add); the
    solver must stop there instead of falling into its bracket/bisection path (regression)."""
    rng = np.random.default_rng(0)
    te = np.array([114, 132, 150, 176, 202.0])
    n = 50000
    t2 = np.exp(rng.uniform(np.log(11), np.log(1900), n))
    k = rng.uniform(1, 9000, n)
    clean = (k[:, None] * np.exp(-te[None, :] / t2[:, None])).astype(np.float32)
    o = hostsim.fit(clean, te, "gaussian", [650, 165], [(0, 10000), (10, 2000)], True)
    rel = np.abs(o["t2"] - t2) / t2
    assert (o["status"] == 0).all() and o["nit"].max() <= 3
    assert rel.max() < 2e-4 and np.median(rel) < 1e-5
