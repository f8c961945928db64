"""Host-side front of the loader (t2fit_host_mask_union_indices / t2fit_host_gather_planes; no GPU involved) against numpy:
``np.sum(mask, axis=3) > 0`` (run_t2mapping.py:384), ``mask[label == 0] = 0`` (:393-400), ``np.where(mask.flatten())[0]``
(:421) and the float32 cast of the masked voxels (:411-412)."""
import ctypes as C

import numpy as np
import pytest

from fetal_t2mapping_b200 import _abi


@pytest.fixture(scope="module")
def lib():
    return _abi.load_library()


def _union(lib, masks, label=None):
    n = masks[0].size
    mo, idx, cnt = np.full(n, 7, np.uint8), np.full(n, -1, np.int64), C.c_int64(-1)
    rc = lib.t2fit_host_mask_union_indices((C.c_void_p * len(masks))(*[m.ctypes.data for m in masks]), len(masks),
                                           _abi.DTYPES[masks[0].dtype.name], label.ctypes.data if label is not None else None,
                                           _abi.DTYPES[label.dtype.name] if label is not None else 0, n, mo.ctypes.data,
                                           idx.ctypes.data, C.byref(cnt))
    assert rc == 0, lib.t2fit_last_error()
    return mo, idx, cnt.value


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("n", [0, 1, 7, 63, 64, 65, 1000, 70001])
def test_union_and_indices_match_numpy(lib, dtype, n):
    rng = np.random.default_rng(n + np.dtype(dtype).itemsize)
    n_masks = 1 + n % 5
    lo = 0 if np.dtype(dtype).kind == "u" else -2                  # signed / float planes: the SUM decides, not "any non-zero"
    masks = [np.ascontiguousarray((rng.integers(lo, 3, n) * (rng.random(n) < 0.3)).astype(dtype)) for _ in range(n_masks)]
    if n > 100:
        for m in masks:
            m[40:90] = 0                                               # a run of empty words, and a run of full ones
            m[200:300] = 1
    want = np.sum(np.stack(masks, axis=-1), axis=-1) > 0 if n else np.zeros(0, bool)
    mo, idx, cnt = _union(lib, masks)
    ref_idx = np.flatnonzero(want)
    assert cnt == ref_idx.size and np.array_equal(mo.astype(bool), want) and set(np.unique(mo)) <= {0, 1}
    assert np.array_equal(idx[:cnt], ref_idx) and (idx[cnt:] == -1).all()          # nothing written past the count
    for ldt in (np.int16, np.uint8, np.float32):
        label = (rng.integers(0, 4, n) * (rng.random(n) < 0.6)).astype(ldt)
        mo, idx, cnt = _union(lib, masks, label)
        w2 = want.copy()
        w2[label == 0] = False
        assert np.array_equal(mo.astype(bool), w2) and np.array_equal(idx[:cnt], np.flatnonzero(w2))


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64])
def test_gather_planes_matches_astype_float32(lib, dtype):
    rng = np.random.default_rng(5)
    n, n_planes = 50021, 5
    if np.dtype(dtype).kind == "f":
        planes = [(rng.standard_normal(n) * 1e3).astype(dtype) for _ in range(n_planes)]
    else:
        info = np.iinfo(dtype)
        planes = [rng.integers(info.min, info.max, n, endpoint=True).astype(dtype) for _ in range(n_planes)]
    m = rng.random(n) < 0.2
    m[1000:3000] = True                                                # long runs (the 8-in-a-row path) and scattered voxels
    m[3000:3500] = False
    idx = np.flatnonzero(m).astype(np.int64)
    for ld in (idx.size, idx.size + 13):
        out = np.full((n_planes, ld), np.float32(-7), np.float32)
        rc = lib.t2fit_host_gather_planes((C.c_void_p * n_planes)(*[p.ctypes.data for p in planes]), n_planes,
                                          _abi.DTYPES[np.dtype(dtype).name], idx.ctypes.data, idx.size, n, out.ctypes.data, ld)
        assert rc == 0, lib.t2fit_last_error()
        for e in range(n_planes):
            assert np.array_equal(out[e, :idx.size], planes[e][idx].astype(np.float32))
            assert (out[e, idx.size:] == -7).all()


def test_host_front_rejects_bad_arguments(lib):
    n = 100
    planes = [np.zeros(n, np.float32)]
    out = np.zeros((1, 4), np.float32)
    ptrs = (C.c_void_p * 1)(planes[0].ctypes.data)
    for bad in (np.array([0, 5, 100, 7], np.int64), np.array([0, -1, 3, 7], np.int64)):
        assert lib.t2fit_host_gather_planes(ptrs, 1, _abi.DTYPES["float32"], bad.ctypes.data, 4, n, out.ctypes.data, 4) == -1
        assert b"outside" in lib.t2fit_last_error()
    ok = np.array([0, 5, 99, 7], np.int64)
    assert lib.t2fit_host_gather_planes(ptrs, 1, _abi.DTYPES["float32"], ok.ctypes.data, 4, n, out.ctypes.data, 3) == -1   # ld < n_fit
    assert lib.t2fit_host_gather_planes(ptrs, 1, 9, ok.ctypes.data, 4, n, out.ctypes.data, 4) == -1                        # dtype code
    cnt = C.c_int64()
    mo, idx = np.zeros(n, np.uint8), np.zeros(n, np.int64)
    assert lib.t2fit_host_mask_union_indices(ptrs, 0, 0, None, 0, n, mo.ctypes.data, idx.ctypes.data, C.byref(cnt)) == -1
    assert lib.t2fit_host_mask_union_indices(ptrs, 1, 9, None, 0, n, mo.ctypes.data, idx.ctypes.data, C.byref(cnt)) == -1
