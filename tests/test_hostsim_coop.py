"""The cooperative (lane-group-per-voxel, shared-memory state) form of the reference-faithful solver
(csrc/t2fit_lbfgsb_coop.cuh) on the lane emulator (tests/hostsim/lane_emu.h), GPU-less:

(1) bit for bit against the serial solver -- same operations in the same order, checked with a build without FMA
    contraction so that expression shapes cannot blur it -- for two group widths and for both lane schedules (lanes run
    forwards / backwards between barriers: a missing barrier changes the result of one of them);
(2) with contraction (as shipped) against the golden fixtures of the unmodified reference, same thresholds as the serial
    solver (tests/test_hostsim_lbfgsb.py);
(3) neither form may depend on what the previous voxel left in memory (scipy zero-initialises the L-BFGS-B workspace for
    every minimize() call; the algorithm reads entries of WN1 it never computed when formk was skipped at an update).
"""
import numpy as np
import pytest

from tests import hostsim
from tests.conftest import assert_lbfgsb_parity, fit_params_of, lbfgsb_parity_report, load_golden

FIELDS = ("x", "fun", "nit", "nfev", "status", "result", "trace_f", "trace_step", "trace_len")


def _run(g, rows, **kw):
    fp = fit_params_of(g)
    return hostsim.lbfgsb(rows, g["te"], g["fit"], g["x0"], g["bounds"], g["prior"], g["norm"], options=fp["options"],
                          trace_cap=24, **kw)


@pytest.mark.parametrize("name,m", [("c2_gaussian_noprior", 250), ("c3_floor_noprior", 900), ("c5_floor_noprior", 250),
                                   ("c3_rician_prior", 200), ("cli3_rician_lf_noprior", 150), ("norm_gaussian", 100),
                                   ("edge_gaussian_rician_noprior", 10 ** 6), ("edge_gaussian_prior", 10 ** 6)])
def test_coop_equals_serial_bit_for_bit(name, m):
    g = load_golden(name)
    rows = g["rows"][:m]
    ref = _run(g, rows, strict=True)
    for lanes, reverse in ((8, False), (8, True), (32, True), (16, False)):
        got = _run(g, rows, strict=True, coop_lanes=lanes, reverse=reverse)
        for k in FIELDS:
            assert np.array_equal(ref[k], got[k], equal_nan=True), (name, lanes, reverse, k,
                                                                     np.flatnonzero(~np.all(np.atleast_2d(ref[k].T == got[k].T), axis=0))[:5])


@pytest.mark.parametrize("name", ["c2_gaussian_noprior", "c3_floor_noprior", "c3_rician_prior", "cli3_floor_hf_prior"])
def test_coop_reproduces_reference_fixtures(name):
    g = load_golden(name)
    o = _run(g, g["rows"], coop_lanes=8)
    rep = lbfgsb_parity_report(o["x"][:, 1], o["nit"], o["status"] == 0, g)
    assert_lbfgsb_parity(rep, name)


def test_result_does_not_depend_on_the_previous_voxel():
    """Voxel 837 of the c3 fixture reads an entry of WN1 that is never computed during its own run (formk skipped at an
    update).  Fitted alone, after its neighbour, or in the whole batch it must give the reference's 15 iterations."""
    g = load_golden("c3_floor_noprior")
    alone = _run(g, g["rows"][837:838], strict=True)
    after = _run(g, g["rows"][836:838], strict=True)
    batch = _run(g, g["rows"][800:840], strict=True)
    assert alone["nit"][0] == after["nit"][1] == batch["nit"][37] == g["ref_nit"][837]
    assert np.array_equal(alone["x"][0], after["x"][1]) and np.array_equal(alone["x"][0], batch["x"][37])
