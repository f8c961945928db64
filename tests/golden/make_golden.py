"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case the reference's own ``set_fit_params`` / ``fit_voxel`` /
``compute_residuals`` (run_t2mapping.py:29-111,120-312; utils/t2map_utils.py:62-89)
are called on seeded synthetic rows; inputs and outputs go to ``<case>.npz``.
Beside the verbatim reference outputs each fixture stores two derived oracles
(SURVEY.md §8(c)): ``tight_*`` (tight-tolerance L-BFGS-B restarted from the
reference's answer -> the *converged voxel* classification) and ``exact_*``
(bounded least-squares minimiser).  Versions used are recorded in the fixture.
"""
import argparse
import contextlib
import copy
import io
import os
import sys
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import fit_oracle as fo                      # noqa: E402
from oracle.ref_loader import load_reference             # noqa: E402
from fetal_t2mapping_b200 import synth                   # noqa: E402


def ref_preset(ref, fit, field):
    ns = argparse.Namespace(gaussian=fit == "gaussian", gaussian_rician=fit == "gaussian_rician",
                            rician=fit == "rician", lf=field == "lf", hf=field == "hf", norm=False)
    return ref.set_fit_params(ns)


def _ref_fit_one(i, ref_fit_voxel, fit, fit_params, te, rows, prior, norm):
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        try:
            p, ok, nit, fun, info = ref_fit_voxel(i, fit, copy.deepcopy(fit_params), te, rows, prior, norm)
            return np.asarray(p, float), bool(ok), int(nit), float(fun), len(info), ""
        except ValueError as e:                          # scipy bounds error kills the reference map
            npar = len(fit_params["initial_guess"])
            return np.full(npar, np.nan), False, -1, np.nan, 0, str(e)


def run_reference_rows(ref, rows, te, fit, fit_params, prior, norm, procs):
    import multiprocessing as mp
    fn = partial(_ref_fit_one, ref_fit_voxel=ref.fit_voxel, fit=fit, fit_params=fit_params, te=te,
                 rows=rows, prior=prior, norm=norm)
    if procs > 1 and rows.shape[0] >= 64:
        with mp.get_context("fork").Pool(procs) as pool:
            out = pool.map(fn, range(rows.shape[0]), chunksize=32)
    else:
        out = [fn(i) for i in range(rows.shape[0])]
    return (np.array([o[0] for o in out]), np.array([o[1] for o in out]),
            np.array([o[2] for o in out], np.int32), np.array([o[3] for o in out]),
            np.array([o[4] for o in out], np.int32), np.array([o[5] for o in out]))


def _exact_one(i, fit, fit_params, te, rows, prior, norm, extra):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fo.fit_voxel_exact(i, fit, fit_params, te, rows, prior, norm, extra_starts=[e[i] for e in extra])


def exact_rows(rows, te, fit, fit_params, prior, norm, extra, procs):
    import multiprocessing as mp
    fn = partial(_exact_one, fit=fit, fit_params=fit_params, te=te, rows=rows, prior=prior, norm=norm,
                 extra=extra)
    with mp.get_context("fork").Pool(procs) as pool:
        out = pool.map(fn, range(rows.shape[0]), chunksize=32)
    return np.array([o[0] for o in out]), np.array([o[1] for o in out])


def sample_rows(cfg, n, scale, rng_seed):
    y, mask, te, truth = synth.make_volume(cfg, scale=scale)
    flat = y.reshape(-1, y.shape[-1])
    idx = np.flatnonzero(mask.reshape(-1))
    rng = np.random.default_rng(rng_seed)
    pick = np.sort(rng.choice(idx, size=min(n, idx.size), replace=False))
    return np.ascontiguousarray(flat[pick]), te, truth["t2"].reshape(-1)[pick], truth["s0"].reshape(-1)[pick]


def build_case(ref, name, rows, te, fit, field, prior, norm, procs, with_exact=True, truth=None):
    fit_r, fp = ref_preset(ref, fit, field)
    assert fit_r == fit
    p, ok, nit, fun, ninfo, err = run_reference_rows(ref, rows, te, fit, fp, prior, norm, procs)
    out = dict(rows=rows, te=te, fit=fit, field=field, prior=prior, norm=norm,
               x0=np.array(fp["initial_guess"], float), bounds=np.array(fp["param_bounds"], float),
               ref_params=p, ref_success=ok, ref_nit=nit, ref_fun=fun, ref_ninfo=ninfo, ref_error=err)
    good = np.isfinite(p).all(axis=1)
    if fit != "rician":
        starts = np.where(good[:, None], p, np.array(fp["initial_guess"], float)[None, :])
        conv, tp = fo.converged_set(rows, te, fit, fp, prior, norm, starts, ok & good, procs=procs)
        out.update(tight_params=tp, converged=conv & good)
        if with_exact:
            ok_rows = np.flatnonzero(good & np.isfinite(rows).all(axis=1))
            ep = np.full_like(p, np.nan)
            ef = np.full(p.shape[0], np.nan)
            if ok_rows.size:
                e_p, e_f = exact_rows(rows[ok_rows], te, fit, fp, prior, norm, [tp[ok_rows]], procs)
                ep[ok_rows], ef[ok_rows] = e_p, e_f
            out.update(exact_params=ep, exact_fun=ef)
    if truth is not None:
        out.update(true_t2=truth[0], true_s0=truth[1])
    import scipy
    out.update(scipy_version=scipy.__version__, numpy_version=np.__version__)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    n = rows.shape[0]
    msg = f"{name}: M={n} success={int(ok.sum())}"
    if "converged" in out:
        msg += f" converged={int(out['converged'].sum())}"
    print(msg, flush=True)


def edge_rows(n_echo):
    """Pathological rows that define the failed / ValueError sets (SURVEY.md §8(a) probes)."""
    base = np.array([700.0, 520.0, 390.0, 240.0, 150.0, 90.0, 60.0, 40.0, 30.0, 20.0, 15.0, 10.0])[:n_echo]
    rows = []
    rows.append(base)                                  # 0 plain decay
    rows.append(np.zeros(n_echo))                      # 1 all zero
    r = base.copy(); r[1] = np.nan; rows.append(r)     # 2 NaN in echo 1
    r = base.copy(); r[0] = np.nan; rows.append(r)     # 3 NaN in echo 0 (ValueError under no_prior)
    r = base.copy(); r[-1] = np.inf; rows.append(r)    # 4 +inf echo
    r = base.copy(); r[0] = np.inf; rows.append(r)     # 5 +inf echo 0
    rows.append(base * 20.0)                           # 6 S(TE0) > 1e4 (ValueError under no_prior)
    rows.append(-base)                                 # 7 all negative
    r = base.copy(); r[1::2] *= -1; rows.append(r)     # 8 alternating sign
    rows.append(np.full(n_echo, 300.0))                # 9 flat (T2 -> upper bound)
    rows.append(base[::-1].copy())                     # 10 increasing (T2 -> upper bound)
    r = np.zeros(n_echo); r[0] = 800.0; rows.append(r)  # 11 instant decay (T2 -> lower bound)
    rows.append(base * 1e-3)                           # 12 tiny
    rows.append(np.full(n_echo, 9999.0))               # 13 near k upper bound
    rows.append(np.full(n_echo, 1e-30))                # 14 denormal-ish
    rows.append(base + 5000.0)                         # 15 large offset
    return np.array(rows, np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ref = load_reference()
    want = lambda n: (not a.only) or (a.only in n)

    if want("c1_gaussian_noprior"):
        rows, te, t2, s0 = sample_rows("c1", 3000, 1.0, 100)
        build_case(ref, "c1_gaussian_noprior", rows, te, "gaussian", "lf", False, False, a.procs, truth=(t2, s0))
    if want("c1_gaussian_prior"):
        rows, te, t2, s0 = sample_rows("c1", 1500, 1.0, 101)
        build_case(ref, "c1_gaussian_prior", rows, te, "gaussian", "lf", True, False, a.procs, truth=(t2, s0))
    if want("c2_gaussian_noprior"):
        rows, te, t2, s0 = sample_rows("c2", 3000, 0.25, 102)
        build_case(ref, "c2_gaussian_noprior", rows, te, "gaussian", "lf", False, False, a.procs, truth=(t2, s0))
    if want("c2_gaussian_hf_prior"):
        rows, te, t2, s0 = sample_rows("c2", 1000, 0.25, 103)
        build_case(ref, "c2_gaussian_hf_prior", rows * np.float32(1.6), te, "gaussian", "hf", True, False,
                   a.procs, truth=(t2, s0 * 1.6))
    if want("c4_gaussian_noprior"):                      # BASELINE config 4: fetal-brain volume, 6 TEs
        rows, te, t2, s0 = sample_rows("c4", 1500, 0.4, 109)
        build_case(ref, "c4_gaussian_noprior", rows, te, "gaussian", "lf", False, False, a.procs, truth=(t2, s0))
    if want("c3_floor_noprior"):
        rows, te, t2, s0 = sample_rows("c3", 1500, 0.2, 104)
        build_case(ref, "c3_floor_noprior", rows, te, "gaussian_rician", "lf", False, False, a.procs, truth=(t2, s0))
    if want("c3_floor_prior"):
        rows, te, t2, s0 = sample_rows("c3", 800, 0.2, 105)
        build_case(ref, "c3_floor_prior", rows, te, "gaussian_rician", "lf", True, False, a.procs, truth=(t2, s0))
    if want("c5_floor_noprior"):
        rows, te, t2, s0 = sample_rows("c5", 800, 0.05, 106)
        build_case(ref, "c5_floor_noprior", rows, te, "gaussian_rician", "lf", False, False, a.procs, truth=(t2, s0))
    if want("c3_rician_prior"):
        rows, te, t2, s0 = sample_rows("c3", 300, 0.2, 107)
        build_case(ref, "c3_rician_prior", rows, te, "rician", "lf", True, False, a.procs, truth=(t2, s0))
    # the CLI's default echo times (3 TEs: [114,202,299] LF / [115,202,299] HF, run_t2mapping.py:540-545) for all three fits,
    # incl. the HF rician preset whose x0 [17,40,0.15] lies outside its bounds and is clipped by scipy (:97-98)
    def cli_rows(n, te, seed, scale, rician):
        rng = np.random.default_rng(seed)
        t2v = rng.uniform(40, 450, n).astype(np.float32)
        s0 = (rng.uniform(500, 1400, n) * scale).astype(np.float32)
        return synth.decay_signal(s0, t2v, te, rng, 14.0 * scale, rician), t2v, s0
    cli = [("cli3_gaussian_lf_noprior", "gaussian", "lf", False, [114.0, 202.0, 299.0], 600, 1.0, False),
           ("cli3_floor_hf_prior", "gaussian_rician", "hf", True, [115.0, 202.0, 299.0], 400, 1.6, True),
           ("cli3_rician_hf_prior", "rician", "hf", True, [115.0, 202.0, 299.0], 300, 1.6, True),
           ("cli3_rician_lf_noprior", "rician", "lf", False, [114.0, 202.0, 299.0], 300, 1.0, True)]
    for i, (nm, fit, field, prior, te_l, n, scale, ric) in enumerate(cli):
        if want(nm):
            te = np.array(te_l)
            rows, t2v, s0 = cli_rows(n, te, 300 + i, scale, ric)
            build_case(ref, nm, rows, te, fit, field, prior, False, a.procs, truth=(t2v, s0))
    for fit, npar in (("gaussian", 3), ("gaussian_rician", 3)):
        for prior in (True, False):
            nm = f"edge_{fit}_{'prior' if prior else 'noprior'}"
            if want(nm):
                build_case(ref, nm, edge_rows(npar), np.array([114.0, 202.0, 299.0]), fit, "lf", prior, False, 1,
                           with_exact=False)
    if want("norm_gaussian"):
        rows, te, t2, s0 = sample_rows("c1", 200, 1.0, 108)
        # --norm has no preset (run_t2mapping.py:107-109) but fit_voxel implements it (:237-240):
        # drive it with the LF gaussian preset and bounds that make sense for a unit-max signal.
        fit_r, fp = ref_preset(ref, "gaussian", "lf")
        fp["initial_guess"] = [1.5, 165]
        fp["param_bounds"] = [(0.5, 20.0), (10, 600)]
        p, ok, nit, fun, ninfo, err = run_reference_rows(ref, rows, te, "gaussian", fp, True, True, a.procs)
        np.savez_compressed(os.path.join(HERE, "norm_gaussian.npz"), rows=rows, te=te, fit="gaussian", field="lf",
                            prior=True, norm=True, x0=np.array(fp["initial_guess"], float),
                            bounds=np.array(fp["param_bounds"], float), ref_params=p, ref_success=ok, ref_nit=nit,
                            ref_fun=fun, ref_ninfo=ninfo, ref_error=err)
        print("norm_gaussian: M=%d success=%d" % (rows.shape[0], ok.sum()))
    if want("block_c1"):
        # whole hot block of process_t2maps on a tiny volume, through the reference's own functions
        y, mask, te, _ = synth.make_volume("c1", scale=0.22)         # 14^3
        mask4 = np.stack([mask] * te.size, axis=-1).astype(np.uint8)
        mask4[..., 1] = 0                                            # per-TE masks differ; union is used (:384)
        for fit in ("gaussian", "gaussian_rician"):
            fit_r, fp = ref_preset(ref, fit, "lf")
            m = np.sum(mask4, axis=3) > 0
            flat = np.reshape(y, (-1, te.size)).astype(np.float32)
            idx, _ = np.where(np.reshape(m, (-1, 1)))
            rows = flat[idx]
            p, ok, nit, fun, ninfo, err = run_reference_rows(ref, rows, te, fit, fp, False, False, a.procs)
            t2_map = np.zeros_like(flat[..., 0]); k_map = np.zeros_like(t2_map)
            s_map = np.zeros_like(t2_map); r_map = np.zeros_like(t2_map)
            t2_map[idx] = p[:, 1].astype(np.float32); k_map[idx] = p[:, 0].astype(np.float32)
            if fit != "gaussian":
                s_map[idx] = p[:, 2].astype(np.float32)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r3 = ref.compute_residuals(flat, te, fit, False, k_map, t2_map, s_map, r_map, idx, m)
            np.savez_compressed(os.path.join(HERE, f"block_c1_{fit}.npz"), t2w=y, mask4=mask4, te=te, fit=fit,
                                t2=t2_map.reshape(m.shape), k=k_map.reshape(m.shape),
                                sigma=s_map.reshape(m.shape), res=r3, mask_indices=idx, ref_success=ok,
                                ref_nit=nit, ref_fun=fun, ref_params=p)
            print(f"block_c1_{fit}: M={idx.size}")
    if want("kat_notebook"):
        # notebooks/20240910_ada_jmri.ipynb cell 15: printed per-TE WM means, 9 TEs; recorded fit on the
        # (unprinted) medians gave x=[369.3,117.6] nit=13 nfev=78.  Loose known-answer test only.
        te = np.array([114, 132, 150, 176, 202, 229, 255, 273, 299], float)
        means = np.array([[141.99, 121.87, 104.77, 83.86, 67.57, 54.28, 44.00, 38.35, 31.66]], np.float32)
        fp = {"initial_guess": [630, 165], "param_bounds": [(float(means[0, 0]), 10000), (10, 600)],
              "solver": "L-BFGS-B", "options": {"ftol": 1e-6, "maxls": 50, "disp": False}}
        p, ok, nit, fun, ninfo, err = run_reference_rows(ref, means, te, "gaussian", fp, True, False, 1)
        np.savez_compressed(os.path.join(HERE, "kat_notebook.npz"), rows=means, te=te, fit="gaussian",
                            x0=np.array([630, 165.0]), bounds=np.array(fp["param_bounds"], float),
                            ref_params=p, ref_success=ok, ref_nit=nit, ref_fun=fun,
                            recorded_x=np.array([369.3, 117.6]), recorded_nit=13)
        print("kat_notebook:", p, nit, fun)


if __name__ == "__main__":
    main()
