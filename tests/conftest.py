import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k: d[k] for k in d.files}
    for k in ("fit", "field"):
        if k in out:
            out[k] = str(out[k])
    for k in ("prior", "norm"):
        if k in out:
            out[k] = bool(out[k])
    return out


def fit_params_of(g):
    """The fit_params dict the fixture was generated with: its x0 / bounds and the preset's scipy options
    (run_t2mapping.py:36-106; every fixture uses the options of its (fit, field) preset)."""
    from fetal_t2mapping_b200 import presets
    opts = presets.preset(g["fit"], g.get("field", "lf") == "lf")[1]["options"]
    return {"initial_guess": [float(v) for v in g["x0"]],
            "param_bounds": [(float(a), float(b)) for a, b in g["bounds"]],
            "solver": "L-BFGS-B", "options": opts}


def lbfgsb_parity_report(t2, nit, success, g):
    """Parity of a reference-faithful (L-BFGS-B) result with a fixture, judged against the reference's own
    reproducibility floor (tests/golden/make_jitter.py): returns a dict of rates."""
    ref, rp = g["ref_params"], g["reproducible"]
    with np.errstate(all="ignore"):
        rel = np.abs(np.asarray(t2, np.float64) - ref[:, 1]) / np.abs(ref[:, 1])
        relj = np.abs(g["jit_params"][:, 1] - ref[:, 1]) / np.abs(ref[:, 1])
    conv = g["converged"] if "converged" in g else np.zeros(rel.shape[0], bool)
    conv_rate = float(np.mean(rel[conv] <= 1e-3)) if conv.sum() >= 20 else None
    return dict(converged=conv_rate, n_converged=int(conv.sum()), jitter_converged=float(np.mean(relj[conv] <= 1e-3)) if conv.sum() >= 20 else None,
                all=float(np.mean(rel <= 1e-3)), jitter_all=float(np.mean(relj <= 1e-3)),
                reproducible=float(np.mean(rel[rp] <= 1e-3)), n_reproducible=int(rp.sum()),
                nit_eq=float(np.mean(np.asarray(nit) == g["ref_nit"])), jitter_nit_eq=float(np.mean(g["jit_nit"] == g["ref_nit"])),
                success_eq=bool(np.array_equal(np.asarray(success, bool), g["ref_success"])))


def assert_lbfgsb_parity(rep, name):
    """Tolerances: identical success sets; relative |dT2| <= 1e-3 on >= 99 % of the CONVERGED voxels and on >= 98.5 % of the voxels whose reference
    result is reproducible under a 1-ulp change of exp (the rest differ between two runs of the reference
    itself; measured 98.97-100 % over the 13 fixtures -- log / i0e also differ in the last ulps, which the
    exp-only jitter does not probe); over ALL voxels no worse than the reference's own jittered rerun by more
    than 1.5 points."""
    print(f"[lbfgsb parity] {name}: T2 within 1e-3 on converged {rep['converged']} (n={rep['n_converged']}; the reference's own "
          f"jittered rerun: {rep['jitter_converged']}), reproducible {rep['reproducible']:.4f} (n={rep['n_reproducible']}), "
          f"all {rep['all']:.4f} (rerun {rep['jitter_all']:.4f}), nit equal {rep['nit_eq']:.4f} (rerun {rep['jitter_nit_eq']:.4f})")
    assert rep["success_eq"], name
    # north_star's sentence, literally: relative |dT2| <= 1e-3 on CONVERGED voxels (reference success and a tight restart from
    # its own answer moves T2 by <= 1e-4; the rician fixtures have no such set: no tight oracle for the NLL).
    # Measured 99.3-100 % (host build) / 98.0-100 % (GPU) over the fixtures that have one; the reference's own jittered rerun
    # reaches 98.0-100 % on it.
    # Threshold: 99 %, or what the reference's own jittered rerun reaches on that set if that is lower, with 1.5 voxels of
    # slack (the converged sets of the loose 3-parameter presets are small: 51-286 voxels, one voxel is 0.35-2 points).
    if rep["converged"] is not None:
        assert rep["converged"] >= min(0.99, rep["jitter_converged"]) - 1.5 / rep["n_converged"], (name, rep)
    assert rep["reproducible"] >= 0.985, (name, rep)
    assert rep["all"] >= rep["jitter_all"] - 0.015, (name, rep)
    assert rep["nit_eq"] >= rep["jitter_nit_eq"] - 0.03, (name, rep)


@pytest.fixture(scope="session")
def gpu_lib():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fetal_t2mapping_b200 as t2
    t2.init(0)
    return t2
