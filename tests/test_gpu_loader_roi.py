"""Rows (f3)/(f4) of the scope table on the GPU: the batched per-TE loader (mask union, --in_vitro_fast label masking,
PLANES layout, overlapped staging; route="auto" = mask part on the host + masked voxels only over PCIe for sparse masks,
route="device" = everything on the GPU) against the one-volume path, and the phantom ROI statistics against numpy."""
import numpy as np
import pytest

from fetal_t2mapping_b200 import synth

pytestmark = pytest.mark.gpu


def _volume(shape, te, seed, mask_dtype=np.uint8, dense=False):
    rng = np.random.default_rng(seed)
    n = int(np.prod(shape))
    t2 = rng.uniform(60, 300, n)
    s0 = rng.uniform(300, 900, n)
    y = (s0[:, None] * np.exp(-te[None, :] / t2[:, None]) + rng.normal(0, 6, (n, te.size))).astype(np.float32)
    t2w = [np.ascontiguousarray(y[:, e].reshape(shape)) for e in range(te.size)]      # per-TE volumes, as read from disk
    base = synth.ellipsoid_mask(shape, [(0.7 if dense else 0.45) * s for s in shape])     # dense: > half of the volume
    masks = []
    for e in range(te.size):                                                          # per-TE masks differ slightly
        m = base.copy()
        m.reshape(-1)[rng.integers(0, n, 20)] ^= True
        masks.append(m.astype(mask_dtype))
    label = np.zeros(shape, np.int16)
    label.reshape(-1)[rng.choice(n, n // 3, replace=False)] = rng.integers(1, 6, n // 3)
    return t2w, masks, label


@pytest.mark.parametrize("route", ["auto", "device"])
@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
def test_series_loader_matches_single_volume_path(gpu_lib, fit, route):
    te = np.array([114.0, 150.0, 202.0, 299.0])
    _, fp = gpu_lib.preset(fit, True)
    shapes = [(12, 14, 10), (20, 18, 16), (12, 14, 10), (9, 7, 11), (20, 18, 16), (12, 14, 10), (11, 13, 9)]
    mask_dt = [np.uint8, np.float32, np.bool_, np.float64, np.int16, np.uint8, np.int32]
    vols = [_volume(s, te, 10 + i, mask_dt[i], dense=i in (2, 5)) for i, s in enumerate(shapes)]
    vols[3] = ([a.astype(np.float64) for a in vols[3][0]], vols[3][1], vols[3][2])    # float64 volumes, as nibabel's get_fdata gives
    vols[4][1][0][...] = -1                                                           # a signed mask plane: the SUM decides (:384)
    assert any(2 * int((np.sum(np.stack(v[1], -1), axis=3) > 0).sum()) > v[1][0].size for v in vols)      # both host routes run
    out = list(gpu_lib.t2map_series(((v[0], v[1]) for v in vols), te, fit, fp, prior=False, solver="fast", depth=2, route=route))
    assert len(out) == len(vols)
    for (t2w, masks, _), r in zip(vols, out):
        stack = np.stack(t2w, axis=-1).astype(np.float32)                             # :385, :411
        mask4 = np.stack(masks, axis=-1)                                              # :383
        ref = gpu_lib.t2map_volume(stack, mask4, te, fit, fp, prior=False, solver="fast")
        union = np.sum(mask4, axis=3) > 0                                             # :384
        assert np.array_equal(r.mask, union) and r.n_fit == int(union.sum()) and r.failed == 0
        for a, b in zip((r.t2, r.k, r.sigma, r.res), ref):
            assert a.dtype == np.float32 and a.shape == union.shape
            assert np.array_equal(a, b)                                               # same kernel, other layout: bit-identical
            assert (a[~union] == 0).all()


def test_series_loader_default_solver_is_the_faithful_one_for_three_parameter_fits(gpu_lib):
    """gaussian_rician / rician through the loader (PLANES layout) with the default solver = L-BFGS-B: identical to the
    one-volume path on the stacked arrays (AoS layout), both fits."""
    te = np.array([114.0, 150.0, 202.0, 299.0])
    vols = [_volume((10, 9, 8), te, 70 + i) for i in range(2)]
    vols = [([np.abs(a) + 1.0 for a in v[0]], v[1], v[2]) for v in vols]      # magnitude data: the Rician NLL takes log(signal)
    for fit in ("gaussian_rician", "rician"):
        _, fp = gpu_lib.preset(fit, True)
        out = list(gpu_lib.t2map_series(((v[0], v[1]) for v in vols), te, fit, fp, prior=True))
        for (t2w, masks, _), r in zip(vols, out):
            ref = gpu_lib.t2map_volume(np.stack(t2w, axis=-1), np.stack(masks, axis=-1), te, fit, fp, prior=True)
            for a, b in zip((r.t2, r.k, r.sigma, r.res), ref):
                assert np.array_equal(a, b)
            assert (r.sigma[r.mask] >= 2.0).all()          # the 3-parameter fits fill the sigma map


@pytest.mark.parametrize("route", ["auto", "device"])
def test_series_loader_in_vitro_fast_label_masking(gpu_lib, route):
    te = np.array([114.0, 202.0, 299.0])
    _, fp = gpu_lib.preset("gaussian", True)
    vols = [_volume((16, 12, 14), te, 40 + i) for i in range(3)]
    out = list(gpu_lib.t2map_series(vols, te, "gaussian", fp, prior=True, fast=True, route=route))
    for (t2w, masks, label), r in zip(vols, out):
        m = np.sum(np.stack(masks, -1), axis=3) > 0
        m[label == 0] = 0                                                             # :393-400
        assert np.array_equal(r.mask, m)
        assert (r.t2[~m] == 0).all() and (r.t2[m] > 0).all()
    # without `fast` the label is ignored
    out2 = list(gpu_lib.t2map_series(vols, te, "gaussian", fp, prior=True, fast=False, route=route))
    assert out2[0].n_fit > out[0].n_fit


@pytest.mark.parametrize("route", ["auto", "device"])
def test_series_loader_bounds_error_aborts_like_the_reference(gpu_lib, route):
    te = np.array([114.0, 202.0, 299.0])
    _, fp = gpu_lib.preset("gaussian", True)
    t2w, masks, _ = _volume((8, 8, 8), te, 3)
    t2w[0][4, 4, 4] = 20000.0                                                         # S(TE0) > 10000 under --no_prior
    for m in masks:
        m[4, 4, 4] = 1
    with pytest.raises(ValueError):
        list(gpu_lib.t2map_series([(t2w, masks)], te, "gaussian", fp, prior=False, route=route))


def test_planes_layout_host_and_device_equal_aos(gpu_lib):
    """T2FIT_LAYOUT_PLANES through the raw C ABI, host and device memory, against the AoS layout."""
    import ctypes as C
    import torch
    from fetal_t2mapping_b200 import _abi
    from fetal_t2mapping_b200.api import _fill_problem
    te = np.array([114.0, 132.0, 150.0, 176.0, 202.0])
    rng = np.random.default_rng(5)
    n = 5000
    y = (rng.uniform(300, 900, n)[:, None] * np.exp(-te[None, :] / rng.uniform(60, 300, n)[:, None]) + rng.normal(0, 6, (n, 5))).astype(np.float32)
    idx = np.sort(rng.choice(n, 1700, replace=False)).astype(np.int64)
    _, fp = gpu_lib.preset("gaussian", True)
    ref = gpu_lib.fit_voxels_batch(y, idx, te, "gaussian", fp, prior=False)
    planes = np.ascontiguousarray(y.T)                                                # [E, N]
    lib = gpu_lib.init()
    for dev in (False, True):
        p, o = _abi.Problem(), _abi.Outputs()
        keep = _fill_problem(p, "gaussian", fp, te, False, False, 0, 0.0, "loglinear", "fast")
        if dev:
            pd_, id_ = torch.from_numpy(planes).cuda(), torch.from_numpy(idx).cuda()
            out = torch.zeros((3, idx.size), dtype=torch.float32, device="cuda")
            p.echoes, p.memory, p.mask_idx = pd_.data_ptr(), _abi.MEM_DEVICE, id_.data_ptr()
            o.t2, o.k, o.res = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
        else:
            out = np.zeros((3, idx.size), np.float32)
            p.echoes, p.memory, p.mask_idx = planes.ctypes.data, _abi.MEM_HOST, idx.ctypes.data
            o.t2, o.k, o.res = out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data
        p.layout, p.ld, p.n_vox, p.n_fit = _abi.LAYOUT_PLANES, n, n, idx.size
        assert lib.t2fit_run(C.byref(p), C.byref(o), None) == 0, lib.t2fit_last_error()
        if dev:
            torch.cuda.synchronize()
            out = out.cpu().numpy()
        assert np.array_equal(out[0], ref.t2) and np.array_equal(out[1], ref.k) and np.array_equal(out[2], ref.res)
        del keep


def test_roi_stats_match_numpy_nanmean_nanstd(gpu_lib, tmp_path):
    rng = np.random.default_rng(0)
    shape = (20, 24, 18)
    t2 = rng.uniform(10, 2000, shape).astype(np.float32)
    k = rng.uniform(100, 3000, shape).astype(np.float32)
    sg = rng.uniform(0, 50, shape).astype(np.float32)
    label = rng.integers(0, 16, shape).astype(np.int16)          # 0 = background, 15 is beyond n_roi
    t2.reshape(-1)[rng.choice(t2.size, 50, replace=False)] = np.nan
    label[label == 7] = 0                                         # an empty ROI: numpy gives NaN
    gt, id = gpu_lib.set_phantom_gt(False)
    assert len(gt) == 14 and id[0] == "T2-1" and gpu_lib.set_phantom_gt(True)[0][0] == 594
    tab = gpu_lib.phantom_roi_table(t2, k, sg, label, id, gt)
    assert list(tab) == ["id", "trueT2", "meanT2", "stdT2", "meanK", "stdK", "meanC", "stdC"]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(len(gt)):                                  # utils/t2map_utils.py:38-45
            sel = label == i + 1
            for m, mean_c, std_c in ((t2, "meanT2", "stdT2"), (k, "meanK", "stdK"), (sg, "meanC", "stdC")):
                em, es = np.nanmean(m[sel].astype(np.float64)), np.nanstd(m[sel].astype(np.float64))
                assert np.isclose(tab[mean_c][i], em, rtol=1e-10, equal_nan=True)
                assert np.isclose(tab[std_c][i], es, rtol=1e-8, equal_nan=True)
    assert np.isnan(tab["meanT2"][6])
    # device tensors in, CSV out
    import torch
    path = tmp_path / "roi.csv"
    tab2 = gpu_lib.save_phantom_csv(torch.from_numpy(t2).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(sg).cuda(),
                                    torch.from_numpy(label.astype(np.int32)).cuda(), id, gt, str(path))
    assert np.allclose(tab2["meanK"], tab["meanK"], equal_nan=True)
    lines = path.read_text().strip().split("\n")
    assert lines[0] == "id,trueT2,meanT2,stdT2,meanK,stdK,meanC,stdC" and len(lines) == 15
    assert lines[1].split(",")[0] == "T2-1" and lines[1].split(",")[1] == "1044"
