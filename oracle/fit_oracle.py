"""numpy/scipy restatement of the reference's per-voxel T2 fit hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA
path and the timed CPU arm of ``bench.py``; never imported by the product.

What is restated (all citations into the reference tree):

* :func:`preset`            <- ``set_fit_params``      run_t2mapping.py:29-111
* :func:`fit_voxel_oracle`  <- ``fit_voxel``           run_t2mapping.py:120-312
* :func:`residual_map`      <- ``compute_residuals``   utils/t2map_utils.py:62-89
* :func:`fit_block_oracle`  <- hot block of ``process_t2maps``
                                                       run_t2mapping.py:383-386,411-461

The arithmetic of the optimiser is NOT in the reference tree: ``fit_voxel``
hands the objective to ``scipy.optimize.minimize(method="L-BFGS-B", jac=False)``
(run_t2mapping.py:261-286), i.e. the third-party L-BFGS-B 3.0 code plus scipy's
2-point finite-difference gradient (``scipy/optimize/_lbfgsb_py.py``,
``_numdiff.py``).  Reference pin: scipy 1.11.3 / numpy 1.26.0
(requirements_frozen.txt:103,144); this image: scipy 1.18.1 / numpy 2.3.  The
oracle calls that dependency with exactly the reference's arguments, so it IS the
reference's arithmetic on this host.  (The product's FP64 solver restates the
published L-BFGS-B algorithm in ``fetal_t2mapping_b200/csrc/t2fit_lbfgsb.cuh``; it is
checked against this oracle and against scipy itself in ``tests/``, never the other
way round.)

Oracle modes:
  ``verbatim``  the reference's own options (ftol/gtol/maxls from the preset).
  ``tight``     same objective / x0 / bounds through L-BFGS-B with
                ftol=1e-15, gtol=1e-10, maxiter=5000.  Optionally restarted
                from a given point (used to classify *converged* voxels).
  ``exact``     bounded least squares (scipy ``least_squares`` TRF, tol 1e-14,
                multi-start): the bounded minimiser of the same objective.

Pinned against the reference run in the build container by
``tests/golden/make_golden.py`` + ``tests/test_oracle_pinned.py``.
"""
from __future__ import annotations

import copy
import io
import contextlib
import os
import warnings
from functools import partial

import numpy as np
from scipy.optimize import least_squares, minimize
from scipy.special import i0e

FITS = ("gaussian", "gaussian_rician", "rician")

# --------------------------------------------------------------------------------------
# presets  (run_t2mapping.py:36-106; order of parameters is (k, T2[, sigma]))
# --------------------------------------------------------------------------------------
_PRESETS = {
    # (fit, field): (x0, bounds, options)
    ("gaussian", "lf"): ([650, 165], [(600, 10000), (10, 600)],
                         {"ftol": 1e-6, "maxls": 50, "disp": False}),                 # :36-46
    ("gaussian_rician", "lf"): ([650, 110, 40], [(550, 10000), (10, 600), (2, 1000)],
                                {"gtol": 1e-2, "ftol": 1e-2, "maxls": 50, "disp": False}),  # :47-58
    ("rician", "lf"): ([650, 110, 40], [(550, 900), (10, 600), (2, 1000)],
                       {"gtol": 1e-2, "ftol": 1e-2, "maxls": 50, "disp": False}),     # :59-70
    ("gaussian", "hf"): ([890, 165], [(850, 30000), (10, 600)],
                         {"ftol": 1e-6, "maxls": 50, "disp": False}),                 # :72-82
    ("gaussian_rician", "hf"): ([890, 110, 40], [(850, 30000), (30, 600), (2, 1000)],
                                {"gtol": 1e-2, "ftol": 1e-2, "maxls": 50, "disp": False}),  # :83-94
    ("rician", "hf"): ([17, 40, 0.15], [(850, 30000), (30, 600), (7, 200)],
                       {"gtol": 1e-2, "ftol": 1e-2, "maxls": 50, "disp": False}),     # :95-106
}


def preset(fit: str, field: str = "lf"):
    """(fit, fit_params) exactly as ``set_fit_params`` builds them (run_t2mapping.py:29-111)."""
    x0, bounds, options = _PRESETS[(fit, field)]
    return fit, {"initial_guess": list(x0), "param_bounds": [tuple(b) for b in bounds],
                 "solver": "L-BFGS-B", "options": dict(options)}


# --------------------------------------------------------------------------------------
# models / objectives  (run_t2mapping.py:129-177)
# --------------------------------------------------------------------------------------
def model_mono(te, k, t2):                      # gauss_model :129-131
    return k * np.exp(-te / t2)


def model_floor(te, k, t2, sigma):              # gauss_rician_model :133-138
    return (k ** 2 * np.exp(-2 * te / t2) + sigma ** 2) ** (1 / 2)


def objective_mono(p, te, y):                   # gauss_obj :141-147  (MEAN squared error)
    r = y - model_mono(te, p[0], p[1])
    return np.sum(r ** 2) / len(y)


def objective_floor(p, te, y):                  # gauss_rician_obj :149-155
    r = y - model_floor(te, p[0], p[1], p[2])
    return np.sum(r ** 2) / len(y)


def objective_rician_nll(p, te, y):             # rician_obj :157-177
    k, t2, sigma = p
    m = model_mono(te, k, t2)
    x = (m * y) / (sigma ** 2)
    ll = np.sum((np.log(y) - np.log(sigma ** 2)) - (y ** 2 + m ** 2) / (2 * sigma ** 2)
                + (np.abs(x) + np.log(i0e(x))))
    return -ll


_OBJECTIVES = {"gaussian": objective_mono, "gaussian_rician": objective_floor,
               "rician": objective_rician_nll}

TIGHT_OPTIONS = {"ftol": 1e-15, "gtol": 1e-10, "maxiter": 5000, "maxfun": 50000, "maxls": 50}


def voxel_signal_and_bounds(voxel, fit_params, reshaped_t2w, prior, norm):
    """Signal row and the bounds ``fit_voxel`` would use for it (run_t2mapping.py:237-245)."""
    row = reshaped_t2w[voxel, :]
    y = row / np.max(row) if norm else row                       # :237-240
    bounds = list(fit_params["param_bounds"])
    if not prior:                                                # :243-245
        bounds[0] = (reshaped_t2w[voxel, 0], 10000)
        bounds[1] = (10, 2000)
    return np.array(y), bounds


def fit_voxel_oracle(voxel, fit, fit_params, TEeffs, reshaped_t2w, prior, norm,
                     mode="verbatim", start=None, trace=True):
    """One voxel, as ``fit_voxel`` does it (run_t2mapping.py:120-312).

    Returns the reference's tuple ``(params, success, nit, fun, iteration_info)``.
    ``iteration_info`` is the callback trace (:180-234): one dict per L-BFGS-B
    iteration with ``f_val``, ``grad_norm`` (always None) and ``step_size``.
    Unlike the reference the caller's ``fit_params`` is not mutated.
    """
    y, bounds = voxel_signal_and_bounds(voxel, fit_params, reshaped_t2w, prior, norm)
    obj = _OBJECTIVES[fit]
    info, prev = [], [None]

    def cb(xk):                                                   # :180-234
        step = np.nan if prev[0] is None else float(np.linalg.norm(xk - prev[0]))
        prev[0] = xk
        info.append({"f_val": obj(xk, TEeffs, y), "grad_norm": None, "step_size": step})

    if mode == "verbatim":
        options = dict(fit_params["options"])
    elif mode == "tight":
        options = dict(TIGHT_OPTIONS)
    else:
        raise ValueError(mode)
    x0 = fit_params["initial_guess"] if start is None else start
    res = minimize(obj, x0, args=(TEeffs, y), method=fit_params["solver"], bounds=bounds,
                   options=options, jac=False, callback=cb if trace else None)   # :260-286
    return res.x, bool(res.success), int(res.nit), res.fun, info          # :289-312


def fit_voxel_exact(voxel, fit, fit_params, TEeffs, reshaped_t2w, prior, norm, extra_starts=()):
    """Bounded least-squares minimiser of the same objective (TRF, tol 1e-14, multi-start).

    Not a reference function: defines the point a converged bounded solver must reach.
    Only for the two least-squares fits.
    """
    y, bounds = voxel_signal_and_bounds(voxel, fit_params, reshaped_t2w, prior, norm)
    y = y.astype(np.float64)
    lb = np.array([b[0] for b in bounds], float)
    ub = np.array([b[1] for b in bounds], float)
    te = np.asarray(TEeffs, float)
    if fit == "gaussian":
        def resid(p):
            return y - model_mono(te, p[0], p[1])
    elif fit == "gaussian_rician":
        def resid(p):
            return y - model_floor(te, p[0], p[1], p[2])
    else:
        raise ValueError(fit)
    starts = [np.clip(np.array(fit_params["initial_guess"], float), lb, ub)]
    starts += [np.clip(np.array(s, float), lb, ub) for s in extra_starts]
    best = None
    for s in starts:
        r = least_squares(resid, s, bounds=(lb, ub), method="trf", xtol=1e-14, ftol=1e-14,
                          gtol=1e-14, max_nfev=2000)
        if best is None or r.cost < best.cost:
            best = r
    return best.x, 2.0 * best.cost / len(y)


# --------------------------------------------------------------------------------------
# residual map  (utils/t2map_utils.py:62-89)
# --------------------------------------------------------------------------------------
def residual_map(reshaped_t2w, TEeffs, fit, norm, k_map, t2_map, sigma_map, res_map,
                 mask_indices, mask):
    """Signed mean residual over the echoes on masked voxels, zeros elsewhere.

    The reference evaluates the model for ALL voxels (:64-71; off-mask T2=0 gives
    exp(-inf)=0 with a divide warning) and stores float32; only ``mask_indices``
    rows reach the output (:84), so only those are evaluated here.
    """
    y = reshaped_t2w[mask_indices]
    k = k_map[mask_indices][:, None]
    t2 = t2_map[mask_indices][:, None]
    te = np.asarray(TEeffs)[None, :]
    with np.errstate(all="ignore"):
        if fit == "gaussian":                                       # :65-67
            pred = (k * np.exp(-te / t2)).astype(reshaped_t2w.dtype)
        else:                                                       # :69-71 (both other fits)
            s = sigma_map[mask_indices][:, None]
            pred = ((k ** 2 * np.exp(-2 * te / t2) + s ** 2) ** (1 / 2)).astype(reshaped_t2w.dtype)
        if norm:                                                    # :74-79
            y = y / np.max(y, axis=1, keepdims=True)
        res_map[mask_indices] = np.sum(y - pred, axis=1) / len(TEeffs)   # :81-84
    return res_map.reshape(mask.shape[:3])                          # :87


# --------------------------------------------------------------------------------------
# the hot block of process_t2maps  (run_t2mapping.py:383-386,411-461)
# --------------------------------------------------------------------------------------
def _pool_init():
    # forked workers: single-threaded BLAS, silence per-voxel "FAIL" prints
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    warnings.filterwarnings("ignore")


def _guarded(voxel, **kw):
    """scipy raises ValueError for lb>ub / non-finite bounds (kills the reference's whole
    pool.map); the batch helpers record it per voxel instead."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return fit_voxel_oracle(voxel, **kw)
        except ValueError:
            npar = len(kw["fit_params"]["initial_guess"])
            return np.full(npar, np.nan), False, -1, np.nan, []


def _fit_one(voxel, **kw):
    return _guarded(voxel, **kw)


def fit_rows_oracle(rows, TEeffs, fit, fit_params, prior, norm, mode="verbatim", procs=1,
                    trace=False, starts=None):
    """Fit every row of ``rows`` f32[M,E] (already gathered at the masked indices).

    ``Pool.map`` over voxel indices as the reference does (run_t2mapping.py:430-443),
    except that the array handed to the workers is the gathered rows, not the whole
    volume (avoids the reference's per-chunk pickling of the full [N,E] array,
    SURVEY.md §6).  Returns (params f64[M,P], success bool[M], nit i32[M], fun f64[M], infos).
    """
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    m = rows.shape[0]
    TEeffs = np.asarray(TEeffs, dtype=np.float64)
    if starts is None:
        fn = partial(_fit_one, fit=fit, fit_params=fit_params, TEeffs=TEeffs, reshaped_t2w=rows,
                     prior=prior, norm=norm, mode=mode, trace=trace)
        args = range(m)
    else:
        fn = partial(_fit_one_start, fit=fit, fit_params=fit_params, TEeffs=TEeffs,
                     reshaped_t2w=rows, prior=prior, norm=norm, mode=mode, trace=trace,
                     starts=np.asarray(starts, float))
        args = range(m)
    if procs <= 1 or m < 64:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = [fn(i) for i in args]
    else:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        with ctx.Pool(processes=procs, initializer=_pool_init) as pool:
            out = pool.map(fn, args, chunksize=max(1, m // (procs * 8)))
    npar = len(fit_params["initial_guess"])
    params = np.array([o[0] for o in out], dtype=np.float64).reshape(m, npar)
    success = np.array([o[1] for o in out], dtype=bool)
    nit = np.array([o[2] for o in out], dtype=np.int32)
    fun = np.array([o[3] for o in out], dtype=np.float64)
    infos = [o[4] for o in out]
    return params, success, nit, fun, infos


def _fit_one_start(voxel, starts=None, **kw):
    return _guarded(voxel, start=starts[voxel], **kw)


def fit_block_oracle(t2w, mask4, TEeffs, fit, fit_params, prior, norm, mode="verbatim", procs=1):
    """Stacked echoes + per-TE masks in, the four float32 maps out.

    Follows run_t2mapping.py:383-386 (mask union), :411-421 (flatten, zero maps,
    ascending C-order masked indices), :430-443 (map over masked voxels),
    :449-458 (scatter with f64->f32 cast), :461 (residual map), :471-473 (reshape).
    ``mask4`` may also be a plain 3-D mask.
    """
    t2w = np.asarray(t2w)
    mask4 = np.asarray(mask4)
    mask = (np.sum(mask4, axis=3) > 0) if mask4.ndim == 4 else (mask4 > 0)
    n_echo = t2w.shape[-1]
    flat = np.reshape(t2w, (-1, n_echo)).astype(np.float32)
    idx = np.flatnonzero(mask.reshape(-1))
    t2_map = np.zeros(flat.shape[0], np.float32)
    k_map = np.zeros_like(t2_map)
    sigma_map = np.zeros_like(t2_map)
    res_map = np.zeros_like(t2_map)
    params, success, nit, fun, infos = fit_rows_oracle(flat[idx], TEeffs, fit, fit_params, prior,
                                                       norm, mode=mode, procs=procs)
    t2_map[idx] = params[:, 1].astype(np.float32)
    k_map[idx] = params[:, 0].astype(np.float32)
    if fit != "gaussian":
        sigma_map[idx] = params[:, 2].astype(np.float32)
    res = residual_map(flat, np.asarray(TEeffs, float), fit, norm, k_map, t2_map, sigma_map,
                       res_map, idx, mask)
    shp = t2w.shape[:3]
    return {"t2": t2_map.reshape(shp), "k": k_map.reshape(shp), "sigma": sigma_map.reshape(shp),
            "res": res, "mask_indices": idx, "success": success, "nit": nit, "fun": fun,
            "params": params, "infos": infos}


def converged_set(rows, TEeffs, fit, fit_params, prior, norm, params, success, procs=1,
                  t2_rtol=1e-4):
    """Classify *converged* voxels (SURVEY.md §7.3 protocol).

    converged := success AND a tight-tolerance L-BFGS-B restart from the reference's
    own answer moves T2 by at most ``t2_rtol`` (relative).  Returns (converged bool[M],
    tight_params f64[M,P]).
    """
    tp, tsucc, _, _, _ = fit_rows_oracle(rows, TEeffs, fit, fit_params, prior, norm, mode="tight",
                                         procs=procs, starts=params)
    with np.errstate(all="ignore"):
        rel = np.abs(tp[:, 1] - params[:, 1]) / np.abs(params[:, 1])
    return success & tsucc & (rel <= t2_rtol), tp
