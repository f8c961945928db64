"""Presets are part of the numerical contract (run_t2mapping.py:36-106): compare with the fixtures'
x0/bounds, which were written by the reference's own set_fit_params."""
import argparse

import numpy as np
import pytest

import fetal_t2mapping_b200 as t2
from tests.conftest import load_golden
from oracle import fit_oracle as fo


@pytest.mark.parametrize("name", ["c1_gaussian_prior", "c2_gaussian_hf_prior", "c3_floor_prior", "c3_rician_prior"])
def test_presets_equal_reference(name):
    g = load_golden(name)
    fit, fp = t2.preset(g["fit"], g["field"] == "lf")
    assert fit == g["fit"]
    assert np.allclose(fp["initial_guess"], g["x0"]) and np.allclose(np.array(fp["param_bounds"], float), g["bounds"])
    assert fp == fo.preset(g["fit"], g["field"])[1]


def test_set_fit_params_signature_and_norm_exit(capsys):
    ns = argparse.Namespace(gaussian=False, gaussian_rician=True, rician=False, lf=False, hf=True, norm=False)
    fit, fp = t2.set_fit_params(ns)
    assert fit == "gaussian_rician" and fp["param_bounds"][1] == (30, 600) and fp["options"]["ftol"] == 1e-2
    ns.norm = True
    with pytest.raises(SystemExit):
        t2.set_fit_params(ns)
