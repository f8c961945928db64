"""Parity of the CUDA path (through the C ABI) with the reference, on the GPU box.

Tolerances (BASELINE.json north_star): relative |dT2| <= 1e-3 on converged voxels, identical
failed-fit sets.  "Converged voxel" (SURVEY.md 7.3): the reference reported success AND a
tight-tolerance L-BFGS-B restart from the reference's own answer moves T2 by <= 1e-4 -- the
classification is stored in the golden fixtures (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from tests.conftest import assert_lbfgsb_parity, fit_params_of, lbfgsb_parity_report, load_golden

pytestmark = pytest.mark.gpu

T2_RTOL = 1e-3          # north_star tolerance


def run_rows(t2, g, device, **kw):
    fp = fit_params_of(g)
    rows = g["rows"]
    if device:
        import torch
        rows = torch.from_numpy(np.ascontiguousarray(rows)).cuda()
    r = t2.fit_voxels_batch(rows, None, g["te"], g["fit"], fp, prior=g["prior"], norm=g["norm"], **kw)
    if device:
        import torch
        torch.cuda.synchronize()
        conv = lambda a: a.cpu().numpy()
        return dict(t2=conv(r.t2), k=conv(r.k), sigma=conv(r.sigma), res=conv(r.res), fun=conv(r.fun),
                    nit=conv(r.nit), status=conv(r.status))
    return dict(t2=r.t2, k=r.k, sigma=r.sigma, res=r.res, fun=r.fun, nit=r.nit, status=r.status)


LB_CASES = ["c1_gaussian_noprior", "c1_gaussian_prior", "c2_gaussian_noprior", "c2_gaussian_hf_prior", "c4_gaussian_noprior", "c3_floor_noprior",
            "c3_floor_prior", "c5_floor_noprior", "c3_rician_prior", "norm_gaussian", "cli3_gaussian_lf_noprior",
            "cli3_floor_hf_prior", "cli3_rician_hf_prior", "cli3_rician_lf_noprior"]


@pytest.mark.parametrize("name", LB_CASES)
def test_lbfgsb_solver_reproduces_the_reference(gpu_lib, name):
    """T2FIT_SOLVER_LBFGSB (the reference's own optimiser restated in FP64) against the reference's outputs for
    all three fit types, including the presets that stop before convergence (ftol = gtol = 1e-2)."""
    g = load_golden(name)
    o = run_rows(gpu_lib, g, True, solver="lbfgsb")
    rep = lbfgsb_parity_report(o["t2"], o["nit"], o["status"] == 0, g)
    assert_lbfgsb_parity(rep, name)
    # k, sigma and the final objective value on the voxels whose iteration count agrees (same trajectory)
    same = (o["nit"] == g["ref_nit"]) & g["reproducible"]
    ref = g["ref_params"]
    rk = np.abs(o["k"][same] - ref[same, 0]) / np.maximum(np.abs(ref[same, 0]), 1.0)
    rf = np.abs(o["fun"][same] - g["ref_fun"][same]) / np.maximum(np.abs(g["ref_fun"][same]), 1e-6)
    if g["fit"] == "gaussian":
        assert np.quantile(rk, 0.99) <= 2e-3 and np.quantile(rf, 0.99) <= 2e-3
    else:   # loose presets (ftol = gtol = 1e-2); with 3 parameters on 3 echoes the objective itself goes to ~0
        assert np.quantile(rk, 0.90) <= 5e-3 and np.quantile(rk, 0.99) <= 0.2
        assert np.quantile(rf, 0.90) <= 2e-2 and np.quantile(rf / np.maximum(1.0, 1.0 / np.maximum(g["ref_fun"][same], 1e-12)), 0.99) <= 0.2
    if g["fit"] != "gaussian":
        # sigma is the loosest direction of these presets (ftol = gtol = 1e-2): 90 % within 1e-2, 99 % within 0.25
        rs = np.abs(o["sigma"][same] - ref[same, 2]) / np.maximum(np.abs(ref[same, 2]), 1.0)
        assert np.quantile(rs, 0.90) <= 1e-2 and np.quantile(rs, 0.99) <= 0.25


@pytest.mark.parametrize("name", ["c2_gaussian_noprior", "c3_floor_noprior", "c3_rician_prior"])
def test_lbfgsb_callback_traces(gpu_lib, name):
    """iteration_info of the reference's callbacks (run_t2mapping.py:180-234): f_val and step_size per iteration,
    NaN step on the first one; host and device memory paths."""
    import torch
    g = load_golden(name)
    nt = g["trace_len"].shape[0]
    fp = fit_params_of(g)
    for dev in (False, True):
        rows = g["rows"][:nt]
        rows = torch.from_numpy(np.ascontiguousarray(rows)).cuda() if dev else rows
        r = gpu_lib.fit_voxels_batch(rows, None, g["te"], g["fit"], fp, prior=g["prior"], norm=g["norm"], solver="lbfgsb",
                                     trace_cap=64)
        infos = r.iteration_infos
        nit = np.asarray(r.nit.cpu() if dev else r.nit)
        assert len(infos) == nt
        checked = 0
        for i in range(nt):
            assert len(infos[i]) == min(nit[i], 64)
            if nit[i] != g["ref_nit"][i] or not g["reproducible"][i]:
                continue
            n = int(g["trace_len"][i])
            assert np.isnan(infos[i][0]["step_size"]) and infos[i][0]["grad_norm"] is None
            f = np.array([d["f_val"] for d in infos[i]])
            st = np.array([d["step_size"] for d in infos[i]])
            # first iterations and the final point agree closely; in between finite-difference noise lets the two
            # trajectories drift by a few percent before they contract to the same answer
            assert np.allclose(f[:min(n, 3)], g["trace_f"][i, :min(n, 3)], rtol=5e-3), (name, i)
            assert np.allclose(f[n - 1], g["trace_f"][i, n - 1], rtol=1e-2), (name, i)
            if n > 1:
                assert np.allclose(st[1], g["trace_step"][i, 1], rtol=2e-2, atol=1e-3), (name, i)
            checked += 1
        assert checked >= nt // 2
        ar = r.as_all_results()
        assert len(ar) == nt and len(ar[0]) == 5 and len(ar[0][4]) == len(infos[0]) and ar[0][2] == nit[0]


def test_rician_failed_set(gpu_lib):
    """rician_obj takes log(signal): any echo <= 0 makes the objective NaN and scipy gives up at the clipped x0
    with success False (SURVEY 8(a) probe)."""
    _, fp = gpu_lib.preset("rician", True)
    te = np.array([114.0, 202.0, 299.0])
    rows = np.array([[700, 390, 150], [700, 0, 150], [700, -5, 150], [np.nan, 390, 150]], np.float32)
    r = gpu_lib.fit_voxels_batch(rows, None, te, "rician", fp, prior=True)
    assert r.solver == "lbfgsb_dense"
    assert list(r.status == 0) == [True, False, False, False]
    assert np.allclose(r.k[1:], 650) and np.allclose(r.t2[1:], 110) and np.allclose(r.sigma[1:], 40)
    with pytest.raises(ValueError):
        gpu_lib.fit_voxels_batch(rows, None, te, "rician", fp, prior=True, solver="fast")


@pytest.mark.parametrize("device", [False, True], ids=["host", "device"])
@pytest.mark.parametrize("name", ["c1_gaussian_noprior", "c1_gaussian_prior", "c2_gaussian_noprior",
                                  "c2_gaussian_hf_prior", "c4_gaussian_noprior"])
def test_gaussian_matches_reference_on_converged_voxels(gpu_lib, name, device):
    g = load_golden(name)
    o = run_rows(gpu_lib, g, device)
    ref, conv, exact = g["ref_params"], g["converged"], g["exact_params"]
    # identical failed-fit sets: the reference succeeded everywhere on these inputs
    assert np.array_equal(o["status"] == 0, g["ref_success"])
    rel = np.abs(o["t2"] - ref[:, 1]) / ref[:, 1]
    assert conv.mean() > 0.97
    assert rel[conv].max() <= T2_RTOL, f"{name}: max rel dT2 on converged voxels {rel[conv].max():.3e}"
    # against the bounded minimiser itself: every voxel, not only the converged ones
    rel_e = np.abs(o["t2"] - exact[:, 1]) / exact[:, 1]
    assert rel_e.max() <= T2_RTOL, f"{name}: max rel dT2 vs exact bounded LSQ {rel_e.max():.3e}"
    # k where it is identifiable (T2 not on its lower bound)
    ident = exact[:, 1] > g["bounds"][1, 0] * 1.01 if g["prior"] else exact[:, 1] > 10.1
    rel_k = np.abs(o["k"] - exact[:, 0]) / np.maximum(np.abs(exact[:, 0]), 1.0)
    assert rel_k[ident].max() <= 2e-3
    # over all voxels the reference itself is only ~98.6-99.9 % within 1e-3 of its own minimiser
    assert (rel <= T2_RTOL).mean() >= 0.98


@pytest.mark.parametrize("name", ["c1_gaussian_noprior", "c2_gaussian_noprior"])
def test_residual_and_fun_match_reference_formulas(gpu_lib, name):
    g = load_golden(name)
    o = run_rows(gpu_lib, g, True)
    te = g["te"][None, :]
    y = g["rows"].astype(np.float64)
    pred = o["k"].astype(np.float64)[:, None] * np.exp(-te / o["t2"].astype(np.float64)[:, None])
    res = (y - pred).sum(1) / te.size                     # utils/t2map_utils.py:81-84
    fun = ((y - pred) ** 2).sum(1) / te.size              # run_t2mapping.py:147
    assert np.abs(o["res"] - res).max() <= 2e-3           # float32 epilogue vs float64 formula, signal ~1e2..1e3
    assert (np.abs(o["fun"] - fun) / np.maximum(fun, 1e-6)).max() <= 1e-3
    conv = g["converged"]
    assert (np.abs(o["fun"][conv] - g["ref_fun"][conv]) / np.maximum(g["ref_fun"][conv], 1e-6)).max() <= 1e-3


@pytest.mark.parametrize("name", ["c3_floor_noprior", "c3_floor_prior", "c5_floor_noprior"])
def test_floor_model_reaches_a_bounded_minimum(gpu_lib, name):
    """3-parameter noise-floor fit, FAST solver (multi-start by default).  The reference stops at ftol=gtol=1e-2
    (run_t2mapping.py:53-54), far from any minimiser, so this solver is NOT the reference's point -- the
    L-BFGS-B solver is (test_lbfgsb_solver_reproduces_the_reference).  What it must be is the bounded minimiser, on ALL
    voxels: (i) failed set identical to the reference's, (ii) cost within 1e-4 of the exact minimum on > 99 % (exact =
    scipy TRF from a grid of starts, tests/golden/make_exact_multistart.py), (iii) T2 within 1e-3 of the exact minimiser on
    >= 95 % (the rest: flat valleys of voxels decayed into the noise floor), (iv) never above the reference's own cost."""
    g = load_golden(name)
    o = run_rows(gpu_lib, g, True, solver="fast")
    assert np.array_equal(o["status"] == 0, g["ref_success"])
    te = g["te"][None, :]
    y = g["rows"].astype(np.float64)

    def mse(k, t2, s):
        m = np.sqrt(k[:, None] ** 2 * np.exp(-2 * te / t2[:, None]) + s[:, None] ** 2)
        return ((y - m) ** 2).mean(1)
    f_mine = mse(o["k"].astype(float), o["t2"].astype(float), o["sigma"].astype(float))
    f_ref = mse(*g["ref_params"].T)
    f_exact = g["exact_fun"]
    assert (f_mine <= f_exact * (1 + 1e-4) + 1e-9).mean() > 0.99
    rel = np.abs(o["t2"] - g["exact_params"][:, 1]) / g["exact_params"][:, 1]
    assert (rel <= T2_RTOL).mean() >= 0.95
    assert (f_mine <= f_ref * (1 + 1e-4) + 1e-6).mean() >= 0.99
    # the single-start variant stays selectable (speed): a local minimiser only
    o1 = run_rows(gpu_lib, g, True, solver="fast", init_mode="loglinear")
    f1 = mse(o1["k"].astype(float), o1["t2"].astype(float), o1["sigma"].astype(float))
    assert (f_mine <= f1 * (1 + 1e-5) + 1e-9).mean() > 0.995


@pytest.mark.parametrize("solver", ["fast", "lbfgsb", "lbfgsb_dense"])
@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
@pytest.mark.parametrize("prior", [True, False], ids=["prior", "noprior"])
def test_edge_cases_failed_sets(gpu_lib, fit, prior, solver):
    g = load_golden(f"edge_{fit}_{'prior' if prior else 'noprior'}")
    fp = fit_params_of(g)
    raises = np.array([len(str(e)) > 0 for e in g["ref_error"]])
    rows = g["rows"]
    if raises.any():
        # scipy raised ValueError inside those voxels -> the reference's whole pool.map aborts
        with pytest.raises(ValueError):
            gpu_lib.fit_voxels_batch(rows, None, g["te"], fit, fp, prior=prior, norm=False, solver=solver)
        rows = rows[~raises]
    keep = ~raises
    r = gpu_lib.fit_voxels_batch(rows, None, g["te"], fit, fp, prior=prior, norm=False, solver=solver)
    # identical failed-fit set
    assert np.array_equal(r.status == 0, g["ref_success"][keep])
    failed = ~g["ref_success"][keep]
    ref = g["ref_params"][keep]
    # failed voxels keep the clipped x0, finite numbers, never NaN
    assert np.allclose(r.t2[failed], ref[failed, 1]) and np.allclose(r.k[failed], ref[failed, 0])
    assert np.isfinite(r.t2).all() and np.isfinite(r.k).all()
    if fit == "gaussian" or solver != "fast":
        ok = g["converged"][keep] & (ref[:, 1] > 10.0)     # T2 on its lower bound: k not identifiable
        rel = np.abs(r.t2[ok] - ref[ok, 1]) / ref[ok, 1]
        if solver == "fast":
            assert rel.max() <= T2_RTOL
        else:       # pathological rows: one ulp of exp can move the reference's own stopping point by percents
            assert np.mean(rel <= T2_RTOL) >= 0.75 and rel.max() <= 5e-2
    if solver != "fast":                                   # same optimiser: same iteration counts on these rows
        assert np.mean(r.nit == g["ref_nit"][keep]) >= 0.7     # 14-16 pathological rows: allow a few flips


def test_norm_path(gpu_lib):
    """--norm (run_t2mapping.py:237-240).  With a unit-max signal the reference's ftol=1e-6 acts on
    max(|f|,1)=1, so the reference itself stops far from its minimiser; compare with the exact bounded
    LSQ minimiser of the same normalised objective instead, and loosely with the reference."""
    from oracle import fit_oracle as fo
    g = load_golden("norm_gaussian")
    o = run_rows(gpu_lib, g, True, solver="fast")
    assert (o["status"] == 0).all()
    fp = fit_params_of(g)
    ex = np.array([fo.fit_voxel_exact(i, "gaussian", fp, g["te"], g["rows"], True, True)[0] for i in range(120)])
    rel_e = np.abs(o["t2"][:120] - ex[:, 1]) / ex[:, 1]
    assert rel_e.max() <= T2_RTOL
    rel = np.abs(o["t2"] - g["ref_params"][:, 1]) / g["ref_params"][:, 1]
    assert np.median(rel) < 5e-3


def test_kat_notebook(gpu_lib):
    g = load_golden("kat_notebook")
    fp = fit_params_of(g)
    r = gpu_lib.fit_voxels_batch(g["rows"], None, g["te"], "gaussian", fp, prior=True)
    assert abs(r.t2[0] - g["ref_params"][0, 1]) / g["ref_params"][0, 1] < 1e-3
    # the notebook's recorded answer was fitted on (unprinted) medians: loose sanity only
    assert abs(r.t2[0] - g["recorded_x"][1]) / g["recorded_x"][1] < 0.05


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
def test_volume_block_matches_reference_block(gpu_lib, fit):
    import torch
    g = load_golden(f"block_c1_{fit}")
    _, fp = gpu_lib.preset(fit, True)
    for dev in (False, True):
        t2w, m4 = g["t2w"], g["mask4"]
        if dev:
            t2w, m4 = torch.from_numpy(t2w).cuda(), torch.from_numpy(m4).cuda()
        maps = gpu_lib.t2map_volume(t2w, m4, g["te"], fit, fp, prior=False)
        if dev:
            torch.cuda.synchronize()
            maps = [m.cpu().numpy() for m in maps]
        t2m, km, sm, rm = maps
        mask = g["mask4"].sum(3) > 0
        assert t2m.shape == mask.shape and t2m.dtype == np.float32
        for m in maps:
            assert (m[~mask] == 0).all()                   # zeros off-mask, never NaN
        assert np.array_equal(np.flatnonzero(t2m.reshape(-1)), g["mask_indices"])
        rel = np.abs(t2m[mask] - g["t2"][mask]) / g["t2"][mask]
        close = (rel <= 1e-4) & (np.abs(km[mask] - g["k"][mask]) <= 1e-4 * np.abs(g["k"][mask])) & \
                (np.abs(sm[mask] - g["sigma"][mask]) <= 1e-3 * np.maximum(g["sigma"][mask], 1.0))
        assert np.abs(rm[mask][close] - g["res"][mask][close]).max() < 0.05
        if fit == "gaussian":
            assert (rel <= T2_RTOL).mean() >= 0.98
            assert (sm == 0).all()
        else:                                              # default solver for this fit: the reference's optimiser
            assert (rel <= T2_RTOL).mean() >= 0.90 and close.mean() >= 0.3
            assert np.abs(sm[mask][close] - g["sigma"][mask][close]).max() < 0.5


def test_host_and_device_paths_bit_identical(gpu_lib):
    import torch
    from fetal_t2mapping_b200 import synth
    y, mask, te, _ = synth.make_volume("c1", scale=0.5)
    flat = y.reshape(-1, te.size)
    idx = np.flatnonzero(mask.reshape(-1))
    for fit in ("gaussian", "gaussian_rician"):
        _, fp = gpu_lib.preset(fit, True)
        a = gpu_lib.fit_voxels_batch(flat, idx, te, fit, fp, prior=False)
        b = gpu_lib.fit_voxels_batch(torch.from_numpy(flat).cuda(), torch.from_numpy(idx).cuda(), te, fit, fp, prior=False)
        torch.cuda.synchronize()
        for f in ("t2", "k", "sigma", "res", "fun", "nit", "status"):
            assert np.array_equal(getattr(a, f), getattr(b, f).cpu().numpy()), f


def test_full_size_c2_properties(gpu_lib):
    """BASELINE config 2 at full size (256^3 x 5 TE, ~1.6 M masked voxels): size-independent properties.
    (a) first-order optimality of every returned point in float64 (projected gradient of the reference's
    objective ~ 0), (b) idempotence: refitting the noise-free model of the fitted parameters returns
    them, (c) zeros off-mask, (d) a seeded 400-voxel sample against the tight oracle."""
    import torch
    from fetal_t2mapping_b200 import synth
    from oracle import fit_oracle as fo
    y, mask, te, _ = synth.make_volume("c2", scale=1.0)
    _, fp = gpu_lib.preset("gaussian", True)
    yd = torch.from_numpy(y).cuda()
    t2m, km, sm, rm = gpu_lib.t2map_volume(yd, torch.from_numpy(mask).cuda(), te, "gaussian", fp, prior=False)
    torch.cuda.synchronize()
    t2m, km, rm = t2m.cpu().numpy(), km.cpu().numpy(), rm.cpu().numpy()
    assert (t2m[~mask] == 0).all() and (km[~mask] == 0).all() and (rm[~mask] == 0).all()
    m = mask.reshape(-1)
    rows = y.reshape(-1, te.size)[m].astype(np.float64)
    k, t2 = km.reshape(-1)[m].astype(np.float64), t2m.reshape(-1)[m].astype(np.float64)
    assert np.isfinite(k).all() and np.isfinite(t2).all() and (t2 >= 10).all() and (t2 <= 2000).all()
    u = np.exp(-te[None, :] / t2[:, None])
    r = rows - k[:, None] * u
    gk = -2 * (r * u).sum(1)
    gt = -2 * (r * k[:, None] * u * te[None, :] / t2[:, None] ** 2).sum(1)
    # scale-free first-order optimality: |g_i| * param_i / (2 * sum y^2), projected on the box
    s = 2 * (rows ** 2).sum(1)
    kl = rows[:, 0]
    pg_k = np.where((k <= kl + 1e-6 * np.abs(kl) + 1e-9) & (gk > 0), 0, np.where((k >= 1e4) & (gk < 0), 0, gk))
    pg_t = np.where((t2 <= 10) & (gt > 0), 0, np.where((t2 >= 2000) & (gt < 0), 0, gt))
    opt = np.maximum(np.abs(pg_k * k), np.abs(pg_t * t2)) / s
    assert np.quantile(opt, 0.999) < 2e-5 and opt.max() < 2e-3, (np.quantile(opt, 0.999), opt.max())
    # (b) idempotence on the noise-free model of the fitted parameters (interior voxels)
    inner = np.flatnonzero((t2 > 11) & (t2 < 1900) & (k > 50) & (k > kl * 1.001) & (k < 9990))[:200000]
    clean = (k[inner, None] * u[inner]).astype(np.float32)
    fr = gpu_lib.fit_voxels_batch(torch.from_numpy(clean).cuda(), None, te, "gaussian",
                                  {"initial_guess": [650, 165], "param_bounds": [(0, 10000), (10, 2000)]}, prior=True)
    torch.cuda.synchronize()
    rel = np.abs(fr.t2.cpu().numpy() - t2[inner]) / t2[inner]
    w = int(np.argmax(rel))
    assert rel.max() < 1e-3 and np.median(rel) < 1e-5, (rel.max(), clean[w], k[inner][w], t2[inner][w],
                                                        float(fr.t2[w]), float(fr.k[w]), int(fr.nit[w]))
    # (d) sample against the tight oracle
    rng = np.random.default_rng(7)
    pick = rng.choice(rows.shape[0], 400, replace=False)
    p, ok, _, _, _ = fo.fit_rows_oracle(rows[pick].astype(np.float32), te, "gaussian", fp, False, False, mode="tight",
                                        procs=4)
    rel = np.abs(t2[pick] - p[:, 1]) / p[:, 1]
    assert ok.all() and (rel <= T2_RTOL).mean() >= 0.995


@pytest.mark.parametrize("fit,solver", [("gaussian", "fast"), ("gaussian_rician", "fast"), ("gaussian_rician", "lbfgsb")])
@pytest.mark.parametrize("shape", [(13, 11, 7), (40, 37, 29), (64, 64, 5)])
@pytest.mark.parametrize("route,density", [("fused", 0.35), ("stream", 0.35), ("fused", 0.02), ("fused", 0.002)])
def test_fused_zero_fill_ragged_volumes(gpu_lib, fit, solver, shape, route, density, monkeypatch):
    """The zero-fill that accompanies a dense fit must zero every unmasked slot (and all of sigma for the
    2-parameter model) whatever the volume size, starting from NaN-poisoned maps, and must never touch a
    masked slot; mask bytes other than 1 count as masked.  Routes: inside the fit launch (a few mask words per fit
    thread; density 0.02 makes that more than one group of words, 0.002 too many, so the side-stream kernel takes
    over) and the side-stream kernel on request."""
    monkeypatch.setenv("T2FIT_FILL", route)
    import ctypes as C
    import torch
    from fetal_t2mapping_b200 import _abi
    from fetal_t2mapping_b200.api import _fill_problem
    rng = np.random.default_rng(sum(shape))
    n = int(np.prod(shape))
    te = np.array([114.0, 150.0, 202.0, 299.0])
    t2 = rng.uniform(60, 300, n)
    y = (rng.uniform(300, 900, n)[:, None] * np.exp(-te[None, :] / t2[:, None]) + rng.normal(0, 5, (n, 4))).astype(np.float32)
    mask = (rng.random(n) < density).astype(np.uint8) * rng.choice([1, 255, 7], n).astype(np.uint8)
    mask[n // 2] = 1
    idx = np.flatnonzero(mask)
    _, fp = gpu_lib.preset(fit, True)
    lib = gpu_lib.init()
    p, o = _abi.Problem(), _abi.Outputs()
    keep = _fill_problem(p, fit, fp, te, False, False, 0, 0.0, "auto", solver)
    yd, idxd, md = torch.from_numpy(y).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(mask).cuda()
    maps = torch.full((4, n), float("nan"), device="cuda")
    p.echoes, p.memory, p.layout, p.mask_idx, p.n_vox, p.n_fit = yd.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS, idxd.data_ptr(), n, idx.size
    o.t2, o.k, o.sigma, o.res, o.dense = maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), maps[3].data_ptr(), 1
    o.zero_fill_mask = md.data_ptr()
    rc = lib.t2fit_run(C.byref(p), C.byref(o), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.t2fit_last_error()
    torch.cuda.synchronize()
    m = maps.cpu().numpy()
    assert not np.isnan(m).any()
    off = mask == 0
    assert (m[:, off] == 0).all()
    ref = gpu_lib.fit_voxels_batch(yd, idxd, te, fit, fp, prior=False, solver=solver)
    torch.cuda.synchronize()
    assert np.array_equal(m[0, idx], ref.t2.cpu().numpy()) and np.array_equal(m[1, idx], ref.k.cpu().numpy())
    assert np.array_equal(m[3, idx], ref.res.cpu().numpy())
    if fit == "gaussian":
        assert (m[2] == 0).all()
    else:
        assert np.array_equal(m[2, idx], ref.sigma.cpu().numpy())
    del keep


@pytest.mark.parametrize("n_echo,layout", [(3, "aos"), (4, "aos"), (6, "soa"), (12, "aos"), (16, "planes")])
@pytest.mark.parametrize("refill", [1, 8, 32])
def test_floor_queue_kernel_equals_one_shot_kernel(gpu_lib, n_echo, layout, refill, monkeypatch):
    """The 3-parameter fast solver runs in a persistent kernel whose lanes pull voxels from a queue; it must return, bit
    for bit, what the one-thread-per-voxel launch returns (same per-voxel arithmetic), whatever the refill threshold,
    for every input layout, with edge rows (NaN / Inf / zero / bounds-invalid) in the mix and a ragged voxel count."""
    import ctypes as C
    import torch
    from fetal_t2mapping_b200 import _abi
    from fetal_t2mapping_b200.api import _fill_problem
    rng = np.random.default_rng(100 * n_echo + refill)
    n = 200011                      # more voxels than the persistent grid has threads (148 SMs x <= 6 x 128): lanes are refilled
    te = np.linspace(100.0, 700.0, n_echo)
    t2v = np.exp(rng.uniform(np.log(10.0), np.log(2000.0), n))
    s = rng.uniform(300, 3000, n)[:, None] * np.exp(-te[None, :] / t2v[:, None])
    y = np.sqrt((s + rng.normal(0, 20, s.shape)) ** 2 + rng.normal(0, 20, s.shape) ** 2).astype(np.float32)
    y[5, 1] = np.nan; y[77, 0] = np.inf; y[123] = 0.0; y[999, 0] = 2.0e4; y[n - 1, n_echo - 1] = -np.inf
    idx = np.unique(np.concatenate([rng.choice(n, 170001, replace=False), [5, 77, 123, 999, n - 1]])).astype(np.int64)
    _, fp = gpu_lib.preset("gaussian_rician", True)
    lib = gpu_lib.init()
    lay = {"aos": _abi.LAYOUT_AOS, "soa": _abi.LAYOUT_SOA, "planes": _abi.LAYOUT_PLANES}[layout]
    if layout == "aos":
        yd, ld = torch.from_numpy(y).cuda(), 0
    elif layout == "planes":
        yd, ld = torch.from_numpy(np.ascontiguousarray(y.T)).cuda(), n
    else:
        yd, ld = torch.from_numpy(np.ascontiguousarray(y[idx].T)).cuda(), idx.size
    idxd = torch.from_numpy(idx).cuda()
    outs = {}
    for kern in ("oneshot", "queue"):
        monkeypatch.setenv("T2FIT_FLOOR_KERNEL", kern)
        monkeypatch.setenv("T2FIT_QUEUE_REFILL", str(refill))
        p, o = _abi.Problem(), _abi.Outputs()
        keep = _fill_problem(p, "gaussian_rician", fp, te, False, False, 0, 0.0, "loglinear", "fast")
        p.echoes, p.memory, p.layout, p.ld = yd.data_ptr(), _abi.MEM_DEVICE, lay, ld
        p.mask_idx, p.n_vox, p.n_fit = (0 if layout == "soa" else idxd.data_ptr()), n, idx.size
        f32 = torch.full((5, idx.size), float("nan"), device="cuda")
        nit = torch.full((idx.size,), -1, dtype=torch.int32, device="cuda")
        st = torch.full((idx.size,), 9, dtype=torch.uint8, device="cuda")
        o.t2, o.k, o.sigma, o.res, o.fun = (f32[j].data_ptr() for j in range(5))
        o.nit, o.status = nit.data_ptr(), st.data_ptr()
        rc = lib.t2fit_run(C.byref(p), C.byref(o), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.t2fit_last_error()
        cnt = (C.c_int64 * 4)()
        assert lib.t2fit_status_counts(torch.cuda.current_stream().cuda_stream, cnt) == 0
        outs[kern] = (f32.cpu().numpy(), nit.cpu().numpy(), st.cpu().numpy(), list(cnt)[1:])
        del keep
    a, b = outs["oneshot"], outs["queue"]
    assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3] == b[3] == [int((a[2] == s_).sum()) for s_ in (1, 2, 3)]
    assert (a[2] == 1).sum() >= 2 and (a[2] == 3).sum() >= 1 and (a[2] == 0).mean() > 0.5 and a[1].max() >= 3, \
        [(a[2] == s_).sum() for s_ in range(4)] + [a[1].max()]


@pytest.mark.parametrize("dtype", ["float64", "int16", "uint16", "int32"])
@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
def test_host_volume_dtypes_are_cast_while_gathering(gpu_lib, dtype, fit):
    """`.astype(np.float32)` of the reference (:411) fused into the host gather: a float64 / integer host volume gives,
    bit for bit, what its float32 copy gives -- through fit_voxels_batch and through t2map_volume."""
    from fetal_t2mapping_b200 import synth
    y, mask, te, _ = synth.make_volume("c1", scale=0.4)
    vol = np.abs(y) + 1.0
    vol = (vol.astype(np.float64) * (1.0 + 1e-9)) if dtype == "float64" else np.round(vol).astype(dtype)
    flat = vol.reshape(-1, te.size)
    idx = np.flatnonzero(mask.reshape(-1))
    _, fp = gpu_lib.preset(fit, True)
    ref = gpu_lib.fit_voxels_batch(flat.astype(np.float32), idx, te, fit, fp, prior=False, solver="fast")
    got = gpu_lib.fit_voxels_batch(flat, idx, te, fit, fp, prior=False, solver="fast")
    for f in ("t2", "k", "sigma", "res", "fun", "nit", "status"):
        assert np.array_equal(getattr(got, f), getattr(ref, f)), f
    all_rows = gpu_lib.fit_voxels_batch(flat, None, te, fit, fp, prior=False, solver="fast")      # no index vector
    assert np.array_equal(all_rows.t2[idx], ref.t2)
    m_ref = gpu_lib.t2map_volume(vol.astype(np.float32), mask, te, fit, fp, prior=False, solver="fast")
    m_got = gpu_lib.t2map_volume(vol, mask, te, fit, fp, prior=False, solver="fast")
    for a, b in zip(m_got, m_ref):
        assert np.array_equal(a, b)
    assert np.array_equal(m_got[0].reshape(-1)[idx], ref.t2)


def test_device_echoes_must_be_float32(gpu_lib):
    import ctypes as C
    import torch
    from fetal_t2mapping_b200 import _abi
    from fetal_t2mapping_b200.api import _fill_problem
    _, fp = gpu_lib.preset("gaussian", True)
    lib = gpu_lib.init()
    p, o = _abi.Problem(), _abi.Outputs()
    keep = _fill_problem(p, "gaussian", fp, [114.0, 202.0, 299.0], False, False, 0, 0.0, "loglinear")
    y = torch.ones((8, 3), device="cuda")
    out = torch.empty(8, device="cuda")
    p.echoes, p.memory, p.layout, p.n_vox, p.n_fit = y.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS, 8, 8
    p.echo_dtype = _abi.ECHO_DTYPES["float64"]
    o.t2 = out.data_ptr()
    assert lib.t2fit_run(C.byref(p), C.byref(o), None) == -1 and b"float32" in lib.t2fit_last_error()
    p.echo_dtype = 77
    p.memory = _abi.MEM_HOST
    assert lib.t2fit_run(C.byref(p), C.byref(o), None) == -1
    del keep


def test_out_of_range_mask_indices_raise(gpu_lib):
    y = np.ones((10, 3), np.float32)
    _, fp = gpu_lib.preset("gaussian", True)
    with pytest.raises(IndexError):
        gpu_lib.fit_voxels_batch(y, np.array([0, 3, 12]), [114.0, 202.0, 299.0], "gaussian", fp)
    with pytest.raises(IndexError):
        gpu_lib.fit_voxels_batch(y, np.array([-1, 3]), [114.0, 202.0, 299.0], "gaussian", fp)
    r = gpu_lib.fit_voxels_batch(y * 700, np.array([0, 3, 9]), [114.0, 202.0, 299.0], "gaussian", fp)   # still usable
    assert r.t2.shape == (3,)


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician", "rician"])
def test_compute_residuals_mirror(gpu_lib, fit):
    """compute_residuals(reshaped_t2w, TEeffs, fit, norm, k_map, t2_map, sigma_map, res_map, mask_indices, mask) as the
    reference calls it (run_t2mapping.py:461), against the oracle's restatement of utils/t2map_utils.py:62-89; 'rician'
    maps are evaluated with the noise-floor model, as the reference does."""
    from oracle import fit_oracle as fo
    rng = np.random.default_rng(1)
    shape = (6, 7, 8)
    n = int(np.prod(shape))
    te = np.array([114.0, 150.0, 202.0, 299.0])
    y = rng.uniform(50, 900, (n, 4)).astype(np.float32)
    mask = rng.random(shape) < 0.4
    idx = np.flatnonzero(mask.reshape(-1))
    k = np.zeros(n, np.float32); t2v = np.zeros(n, np.float32); sg = np.zeros(n, np.float32)
    k[idx] = rng.uniform(300, 900, idx.size); t2v[idx] = rng.uniform(40, 400, idx.size); sg[idx] = rng.uniform(2, 60, idx.size)
    for norm in (False, True):
        got = gpu_lib.compute_residuals(y, te, fit, norm, k, t2v, sg, np.zeros(n, np.float32), idx, mask)
        want = fo.residual_map(y, te, fit, norm, k, t2v, sg, np.zeros(n, np.float32), idx, mask)
        assert got.shape == shape and (got[~mask] == 0).all()
        assert np.allclose(got, want, rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("fit,solver", [("gaussian", "fast"), ("gaussian_rician", "lbfgsb")])
def test_page_locked_input_needs_no_staging(gpu_lib, fit, solver, monkeypatch):
    """Host call with page-locked input: ONE kernel reads the masked rows from host memory and writes the results back
    (run_host_mapped); bit-identical to the staged pipeline, pinned or pageable index vector, out-of-range indices raise."""
    from fetal_t2mapping_b200 import synth
    y, mask, te, _ = synth.make_volume("c1", scale=0.4)
    flat = np.abs(y.reshape(-1, te.size)) + 1.0
    idx = np.flatnonzero(mask.reshape(-1))
    _, fp = gpu_lib.preset(fit, True)
    monkeypatch.setenv("T2FIT_HOST_IN", "staged")
    ref = gpu_lib.fit_voxels_batch(flat, idx, te, fit, fp, prior=False, solver=solver)
    monkeypatch.setenv("T2FIT_HOST_IN", "mapped")
    flat_p = gpu_lib.pinned_array(None, like=flat)
    for ix in (idx, gpu_lib.pinned_array(None, like=idx), None):
        r = gpu_lib.fit_voxels_batch(flat_p, ix, te, fit, fp, prior=False, solver=solver)
        sel = slice(None) if ix is not None else idx
        for f in ("t2", "k", "sigma", "res", "fun", "nit", "status"):
            assert np.array_equal(getattr(r, f)[sel], getattr(ref, f)), f
    with pytest.raises(IndexError):      # checked by the kernel itself (no host pass over the index vector)
        gpu_lib.fit_voxels_batch(flat_p, np.array([0, flat.shape[0]]), te, fit, fp, prior=False, solver=solver)
    with pytest.raises(IndexError):
        gpu_lib.fit_voxels_batch(flat_p, np.array([-1, 3]), te, fit, fp, prior=False, solver=solver)
    r = gpu_lib.fit_voxels_batch(flat_p, idx, te, fit, fp, prior=False, solver=solver)                      # still usable
    assert np.array_equal(r.t2, ref.t2) and np.array_equal(r.status, ref.status)
