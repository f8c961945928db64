"""ABI v4 on the GPU: per-call status counters, range-checked device index vectors, int32 mask indices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _c1(scale=0.3):
    from fetal_t2mapping_b200 import synth
    y, mask, te, _ = synth.make_volume("c1", scale=scale)
    return y, mask, te, np.ascontiguousarray(y.reshape(-1, te.size)), np.flatnonzero(mask.reshape(-1))


def test_status_counts_do_not_leak_between_calls(gpu_lib):
    """A --no_prior volume with a voxel above the k bound makes the loader raise (as scipy does); a clean device call right
    after it must not inherit that count (round 1: one process-global counter nobody cleared)."""
    import torch
    y, mask, te, flat, idx = _c1()
    _, fp = gpu_lib.preset("gaussian", True)
    bad = y.copy()
    z0, y0, x0 = np.argwhere(mask)[5]
    bad[z0, y0, x0, 0] = 2.0e4                                   # S(TE0) > 10000 under --no_prior -> scipy's ValueError
    vols = [([np.ascontiguousarray(bad[..., e]) for e in range(te.size)], [mask.astype(np.uint8)] * te.size)]
    with pytest.raises(ValueError, match="upper bound"):
        list(gpu_lib.t2map_series(vols, te, "gaussian", fp, prior=False))
    # bad-bounds voxels through fit_voxels_into (no check by the callee) ...
    from fetal_t2mapping_b200.api import check_counts, fit_voxels_into
    bflat = torch.from_numpy(np.ascontiguousarray(bad.reshape(-1, te.size))).cuda()
    idxd = torch.from_numpy(idx).cuda()
    out = {k: torch.empty(idx.size, device="cuda") for k in ("t2", "k", "res")}
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    fit_voxels_into(bflat, idxd, te, "gaussian", fp, False, False, {k: v.data_ptr() for k, v in out.items()}, counts=cnt)
    with pytest.raises(ValueError, match="upper bound"):
        check_counts(cnt)
    # ... and a clean call afterwards sees only its own voxels
    r = gpu_lib.fit_voxels_batch(torch.from_numpy(flat).cuda(), idxd, te, "gaussian", fp, prior=False, norm=False)
    assert r.status_count[3] == 0 and r.status_count[0] == idx.size - r.status_count[1] - r.status_count[2]
    maps = gpu_lib.t2map_volume(torch.from_numpy(y).cuda(), torch.from_numpy(mask).cuda(), te, "gaussian", fp, prior=False)
    assert maps[0].shape == mask.shape


@pytest.mark.parametrize("solver", ["fast", "lbfgsb", "lbfgsb_dense"])
def test_device_index_vector_is_range_checked(gpu_lib, solver):
    """A caller-supplied device index tensor with entries outside [0, n_vox): IndexError, as numpy's fancy indexing in the
    reference (:237), instead of out-of-bounds reads / writes."""
    import torch
    _, _, te, flat, idx = _c1()
    _, fp = gpu_lib.preset("gaussian", True)
    yd = torch.from_numpy(flat).cuda()
    for bad in (flat.shape[0], flat.shape[0] + 12345, -1):
        ib = idx.copy()
        ib[7] = bad
        with pytest.raises(IndexError):
            gpu_lib.fit_voxels_batch(yd, torch.from_numpy(ib).cuda(), te, "gaussian", fp, prior=False, norm=False, solver=solver)
    r = gpu_lib.fit_voxels_batch(yd, torch.from_numpy(idx).cuda(), te, "gaussian", fp, prior=False, norm=False, solver=solver)
    assert int((r.status != 0).sum()) == 0


@pytest.mark.parametrize("fit,solver", [("gaussian", "fast"), ("gaussian_rician", "fast"), ("gaussian_rician", "lbfgsb"), ("gaussian_rician", "lbfgsb_dense")])
def test_int32_mask_indices_equal_int64(gpu_lib, fit, solver):
    """mask_indices as int32 (half the index bytes over PCIe): device tensors, pageable numpy (staged) and page-locked
    numpy (the kernel reads the index vector in place) give exactly what int64 gives."""
    import torch
    _, _, te, flat, idx = _c1()
    _, fp = gpu_lib.preset(fit, True)
    kw = dict(prior=False, norm=False, solver=solver)
    ref = gpu_lib.fit_voxels_batch(flat, idx, te, fit, fp, **kw)
    i32 = idx.astype(np.int32)
    cases = {"pageable": (flat, i32), "pinned": (gpu_lib.pinned_array(None, like=flat), gpu_lib.pinned_array(None, like=i32))}
    for name, (a, i) in cases.items():
        r = gpu_lib.fit_voxels_batch(a, i, te, fit, fp, **kw)
        assert np.array_equal(r.t2, ref.t2) and np.array_equal(r.res, ref.res) and np.array_equal(r.status, ref.status), name
    rd = gpu_lib.fit_voxels_batch(torch.from_numpy(flat).cuda(), torch.from_numpy(i32).cuda(), te, fit, fp, **kw)
    assert np.array_equal(rd.t2.cpu().numpy(), ref.t2) and np.array_equal(rd.nit.cpu().numpy(), ref.nit)
    with pytest.raises(IndexError):
        bad = i32.copy(); bad[3] = flat.shape[0]
        gpu_lib.fit_voxels_batch(cases["pinned"][0], gpu_lib.pinned_array(None, like=bad), te, fit, fp, **kw)
    # want= limits the outputs that cross the bus
    r = gpu_lib.fit_voxels_batch(cases["pinned"][0], cases["pinned"][1], te, fit, fp, want=("status",), **kw)
    assert r.fun is None and r.nit is None and np.array_equal(r.t2, ref.t2)
