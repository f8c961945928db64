"""Strengthen the ``exact`` oracle of the 3-parameter fixtures: bounded least-squares minimiser (scipy TRF, tol 1e-14)
from a GRID of start points instead of two.

``exact_params`` / ``exact_fun`` are not outputs of the reference: they define the point a converged bounded solver must
reach (oracle/fit_oracle.py: fit_voxel_exact).  The noise-floor objective has several local minima on voxels whose
signal has decayed into the floor (T2 on its lower bound with sigma carrying the signal, T2 on its upper bound with k
carrying it, the decaying solution in between); with the two starts of make_golden.py (preset x0, the reference's tightened
answer) the oracle missed the global one on ~4 % of the voxels of c3_floor_noprior / c5_floor_noprior (the CUDA multi-start
solver found lower costs there).  This script re-solves every voxel from the previous answer plus a grid of 5 T2 values x 2
sigma values (k from the linear least-squares fit at that T2) and keeps the lowest cost; the fixture files are updated in
place (all other arrays untouched).  Run:  python tests/golden/make_exact_multistart.py [--procs N]
"""
import argparse
import os
import sys
from functools import partial

import numpy as np
from scipy.optimize import least_squares

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["c3_floor_noprior", "c3_floor_prior", "c5_floor_noprior", "cli3_floor_hf_prior"]


def _one(i, rows, te, bounds, prior, prev):
    y = rows[i].astype(np.float64)
    lb, ub = bounds[:, 0].copy(), bounds[:, 1].copy()
    if not prior:                                   # run_t2mapping.py:243-245
        lb[0], ub[0] = float(rows[i, 0]), 10000.0
        lb[1], ub[1] = 10.0, 2000.0
    if not np.isfinite(y).all() or not (lb <= ub).all():
        return prev[i], np.nan

    def resid(p):
        return y - np.sqrt(p[0] ** 2 * np.exp(-2.0 * te / p[1]) + p[2] ** 2)
    starts = [np.clip(prev[i], lb, ub)] if np.isfinite(prev[i]).all() else []
    rms = float(np.sqrt(np.mean(y * y)))
    for t2 in lb[1] * (ub[1] / lb[1]) ** np.linspace(0.0, 1.0, 5):
        u = np.exp(-te / t2)
        k = float(np.dot(y, u) / max(np.dot(u, u), 1e-300))
        for s in (lb[2], rms):
            starts.append(np.clip(np.array([k, t2, s]), lb, ub))
    best = None
    for s in starts:
        r = least_squares(resid, s, bounds=(lb, ub), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=2000)
        if best is None or r.cost < best.cost:
            best = r
    return best.x, 2.0 * best.cost / len(y)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import multiprocessing as mp
    for name in NAMES:
        if a.only and name not in a.only.split(","):
            continue
        path = os.path.join(HERE, name + ".npz")
        d = dict(np.load(path))
        rows, te = d["rows"], d["te"].astype(float)
        prev, prev_f = d["exact_params"], d["exact_fun"]
        fn = partial(_one, rows=rows, te=te, bounds=d["bounds"].astype(float), prior=bool(d["prior"]), prev=prev)
        with mp.Pool(a.procs) as pool:
            out = pool.map(fn, range(rows.shape[0]), chunksize=16)
        ep = np.array([o[0] for o in out])
        ef = np.array([o[1] for o in out])
        keep_old = ~np.isfinite(ef) | (prev_f <= ef)            # never worse than before
        ep[keep_old], ef[keep_old] = prev[keep_old], prev_f[keep_old]
        lower = np.isfinite(ef) & (ef < prev_f * (1 - 1e-6))
        print(f"{name}: {rows.shape[0]} voxels, lower minimum found on {int(lower.sum())} ({100 * lower.mean():.2f} %), "
              f"largest relative drop {np.nanmax((prev_f - ef) / np.maximum(prev_f, 1e-300)):.3g}", flush=True)
        d["exact_params"], d["exact_fun"] = ep, ef
        d["exact_multistart"] = np.array(True)
        np.savez_compressed(path, **d)


if __name__ == "__main__":
    sys.exit(main())
