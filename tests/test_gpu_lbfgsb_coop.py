"""The cooperative L-BFGS-B kernel (lbfgsb_coop_kernel: one voxel per group of 8 / 16 / 32 lanes, optimiser state in shared
memory) on the GPU: against the one-thread-per-voxel kernel (same operations in the same order -> the same results) and
against the golden fixtures of the unmodified reference (same thresholds as the thread kernel, tests/conftest.py)."""
import os

import numpy as np
import pytest

from tests.conftest import assert_lbfgsb_parity, fit_params_of, lbfgsb_parity_report, load_golden

pytestmark = pytest.mark.gpu


def _fit(t2, g, kernel, device=True, trace_cap=0):
    import torch
    fp = fit_params_of(g)
    old = os.environ.get("T2FIT_LB_KERNEL")
    os.environ["T2FIT_LB_KERNEL"] = kernel
    try:
        rows = torch.from_numpy(g["rows"]).cuda() if device else g["rows"]
        r = t2.fit_voxels_batch(rows, None, g["te"], g["fit"], fp, prior=g["prior"], norm=g["norm"], solver="lbfgsb",
                                trace_cap=trace_cap)
        if device:
            torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("T2FIT_LB_KERNEL", None)
        else:
            os.environ["T2FIT_LB_KERNEL"] = old
    host = lambda a: None if a is None else (a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a))
    return {k: host(getattr(r, k)) for k in ("t2", "k", "sigma", "res", "fun", "nit", "status", "trace_f", "trace_step", "trace_len")}


@pytest.mark.parametrize("name", ["c2_gaussian_noprior", "c3_floor_noprior", "c5_floor_noprior", "c3_rician_prior",
                                  "cli3_rician_lf_noprior", "norm_gaussian", "c1_gaussian_prior"])
@pytest.mark.parametrize("kernel", ["coop8", "coop16", "coop32"])
def test_coop_kernel_equals_thread_kernel(gpu_lib, name, kernel):
    g = load_golden(name)
    a = _fit(gpu_lib, g, "thread", trace_cap=16)
    b = _fit(gpu_lib, g, kernel, trace_cap=16)
    assert np.array_equal(a["status"], b["status"])
    same = (a["nit"] == b["nit"]) & (a["t2"] == b["t2"]) & (a["k"] == b["k"]) & (a["fun"] == b["fun"]) & (a["res"] == b["res"])
    # same operations in the same order; what is left to differ is how nvcc contracts a*b+c in the two translation forms
    assert same.mean() >= 0.995, (name, kernel, float(same.mean()))
    assert np.array_equal(a["trace_len"][same], b["trace_len"][same])
    assert np.array_equal(a["trace_f"][same], b["trace_f"][same], equal_nan=True)


@pytest.mark.parametrize("name", ["c1_gaussian_noprior", "c2_gaussian_noprior", "c2_gaussian_hf_prior", "c4_gaussian_noprior", "c3_floor_noprior", "c3_floor_prior",
                                  "c5_floor_noprior", "c3_rician_prior", "cli3_gaussian_lf_noprior", "cli3_floor_hf_prior",
                                  "cli3_rician_hf_prior", "cli3_rician_lf_noprior"])
def test_coop_kernel_reproduces_reference_fixtures(gpu_lib, name):
    g = load_golden(name)
    r = _fit(gpu_lib, g, "coop8")
    rep = lbfgsb_parity_report(r["t2"], r["nit"], r["status"] == 0, g)
    assert_lbfgsb_parity(rep, name)


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
def test_coop_kernel_edge_rows_and_host_path(gpu_lib, fit):
    """NaN / Inf / zero / negative rows through the cooperative kernel on the host path; scipy's ValueError rows (status 3)
    abort the call as the reference's pool.map does."""
    g = load_golden(f"edge_{fit}_noprior")
    raises = np.array([len(str(e)) > 0 for e in g["ref_error"]])
    assert raises.any()
    with pytest.raises(ValueError):
        _fit(gpu_lib, g, "coop8", device=False)
    with pytest.raises(ValueError):
        _fit(gpu_lib, g, "coop8", device=True)
    g = dict(g, rows=np.ascontiguousarray(g["rows"][~raises]))
    a = _fit(gpu_lib, g, "thread", device=False)
    b = _fit(gpu_lib, g, "coop8", device=False)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(b["status"] == 0, g["ref_success"][~raises])
    assert np.array_equal(a["t2"], b["t2"], equal_nan=True) and np.array_equal(a["nit"], b["nit"])


def test_coop_kernel_many_voxels_refill(gpu_lib):
    """170 k voxels: every group is refilled from the queue hundreds of times; compact and dense outputs agree with the
    thread kernel."""
    import torch
    from fetal_t2mapping_b200 import synth
    y, mask, te, _ = synth.make_volume("c2", scale=0.47)
    flat = torch.from_numpy(np.ascontiguousarray(y.reshape(-1, te.size))).cuda()
    idx = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).cuda()
    _, fp = gpu_lib.preset("gaussian", True)
    out = {}
    for kernel in ("thread", "coop8"):
        os.environ["T2FIT_LB_KERNEL"] = kernel
        try:
            r = gpu_lib.fit_voxels_batch(flat, idx, te, "gaussian", fp, prior=False, norm=False, solver="lbfgsb")
            torch.cuda.synchronize()
        finally:
            os.environ.pop("T2FIT_LB_KERNEL", None)
        out[kernel] = r
    a, b = out["thread"], out["coop8"]
    assert idx.numel() > 150000
    assert torch.equal(a.status, b.status)
    same = (a.t2 == b.t2) & (a.nit == b.nit) & (a.res == b.res)
    assert float(same.float().mean()) >= 0.995
