"""Fit presets: the numerical contract of ``set_fit_params`` (run_t2mapping.py:29-111).

Initial guess, bounds and scipy options per (noise model x field strength), parameter
order (k, T2[, sigma]).  The options are kept in the returned dict so existing callers
can pass the reference's ``fit_params`` straight through; the CUDA solver converges to
the bounded minimiser and does not use scipy's ftol/gtol/maxls.
"""
from __future__ import annotations

import copy

__all__ = ["set_fit_params", "preset", "NO_PRIOR_K_UB", "NO_PRIOR_T2_BOUNDS"]

# --no_prior per-voxel override (run_t2mapping.py:243-245)
NO_PRIOR_K_UB = 10000.0
NO_PRIOR_T2_BOUNDS = (10.0, 2000.0)

_LOOSE = {"gtol": 1e-2, "ftol": 1e-2, "maxls": 50, "disp": False}
_GAUSS = {"ftol": 1e-6, "maxls": 50, "disp": False}

_TABLE = {
    ("gaussian", True): ([650, 165], [(600, 10000), (10, 600)], _GAUSS),                            # :36-46
    ("gaussian_rician", True): ([650, 110, 40], [(550, 10000), (10, 600), (2, 1000)], _LOOSE),      # :47-58
    ("rician", True): ([650, 110, 40], [(550, 900), (10, 600), (2, 1000)], _LOOSE),                 # :59-70
    ("gaussian", False): ([890, 165], [(850, 30000), (10, 600)], _GAUSS),                           # :72-82
    ("gaussian_rician", False): ([890, 110, 40], [(850, 30000), (30, 600), (2, 1000)], _LOOSE),     # :83-94
    ("rician", False): ([17, 40, 0.15], [(850, 30000), (30, 600), (7, 200)], _LOOSE),               # :95-106
}


def preset(fit: str, low_field: bool = True):
    """``(fit, fit_params)`` for a noise model and field strength."""
    x0, bounds, opts = _TABLE[(fit, bool(low_field))]
    return fit, {"initial_guess": list(x0), "param_bounds": [tuple(b) for b in bounds],
                 "solver": "L-BFGS-B", "options": copy.deepcopy(opts)}


def set_fit_params(args):
    """Same call as the reference: ``args`` carries the CLI flags ``gaussian``,
    ``gaussian_rician``, ``rician``, ``lf``, ``hf``, ``norm`` (run_t2mapping.py:483-518).
    ``--norm`` has no preset in the reference (it prints an error and exits, :107-109);
    here that is a ``SystemExit(1)`` as well."""
    fit = "gaussian" if getattr(args, "gaussian", False) else \
        "gaussian_rician" if getattr(args, "gaussian_rician", False) else \
        "rician" if getattr(args, "rician", False) else None
    field = True if getattr(args, "lf", False) else False if getattr(args, "hf", False) else None
    if fit is None or field is None or getattr(args, "norm", False):
        print("Error: Normalization is set to true though no parameters where defined yet. "
              "Please modify set_fit_params to manage.")
        raise SystemExit(1)
    return preset(fit, field)
