"""Phantom ROI statistics on the device: the tail of ``process_t2maps`` for in-vitro data.

Mirrors ``set_phantom_gt`` (run_t2mapping.py:14-27) and ``save_phantom_csv``
(utils/t2map_utils.py:30-59): per-label NaN-skipping mean and population standard deviation of the
T2 / k / sigma maps, written as the same CSV.  The reduction runs as a segmented two-pass CUDA
reduction (``t2fit_roi_stats``); numpy inputs are staged through torch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .api import _is_torch, _state, init

__all__ = ["set_phantom_gt", "roi_stats", "phantom_roi_table", "save_phantom_csv"]


def set_phantom_gt(low_field):
    """NMR ground-truth T2 of the phantom spheres and their ids, returned as ``(gt, id)`` like the reference
    (run_t2mapping.py:14-27).  NB the reference's caller unpacks them swapped (``id, gt = set_phantom_gt(...)``,
    :477); :func:`phantom_roi_table` takes them positionally in the order ``save_phantom_csv`` does."""
    if low_field:
        gt = [594, 416, 284, 221, 167, 122, 80, 53, 41]
        id = ["T2-3", "T2-4", "T2-5", "T2-6", "T2-7", "T2-8", "T2-9", "T2-10", "T2-11"]
    else:
        gt = [1044, 624, 428, 258, 186, 137, 90, 63, 44, 27, 19, 15, 10, 8]
        id = ["T2-1", "T2-2", "T2-3", "T2-4", "T2-5", "T2-6", "T2-7", "T2-8", "T2-9", "T2-10", "T2-11", "T2-12",
              "T2-13", "T2-14"]
    return gt, id


def roi_stats(maps, label, n_roi):
    """``(mean, std, count)`` arrays ``[len(maps), n_roi]``: np.nanmean / np.nanstd of every map over
    ``label == i + 1`` (utils/t2map_utils.py:39-45).  ``maps``: float32 arrays or CUDA tensors of one shape."""
    import torch
    lib = init()
    dev = torch.device("cuda", _state["device"])

    def to_dev(a, dt):
        t = a if _is_torch(a) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=dev, dtype=dt).contiguous().reshape(-1)
    md = [to_dev(m, torch.float32) for m in maps]
    lab = to_dev(label, torch.int32)
    n_vox = lab.numel()
    for m in md:
        if m.numel() != n_vox:
            raise ValueError("maps and label must have the same number of voxels")
    ptrs = (C.c_void_p * len(md))(*[m.data_ptr() for m in md])
    mean = np.zeros((len(md), n_roi), np.float64)
    std = np.zeros_like(mean)
    cnt = np.zeros((len(md), n_roi), np.int64)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _abi.check(lib, lib.t2fit_roi_stats(ptrs, len(md), lab.data_ptr(), n_vox, int(n_roi),
                                        mean.ctypes.data_as(C.POINTER(C.c_double)), std.ctypes.data_as(C.POINTER(C.c_double)),
                                        cnt.ctypes.data_as(C.POINTER(C.c_int64)), stream), "t2fit_roi_stats")
    return mean, std, cnt


def phantom_roi_table(t2_map, k_map, sigma_map, label, id, gt):
    """The DataFrame of ``save_phantom_csv`` as an ordered dict of columns (same names, same order:
    id, trueT2, meanT2, stdT2, meanK, stdK, meanC, stdC); ``n_roi = len(gt)`` as in the reference."""
    n_roi = len(gt)
    mean, std, _ = roi_stats([t2_map, k_map, sigma_map], label, n_roi)
    return {"id": list(id), "trueT2": list(gt), "meanT2": mean[0], "stdT2": std[0], "meanK": mean[1], "stdK": std[1],
            "meanC": mean[2], "stdC": std[2]}


def save_phantom_csv(t2_map, k_map, sigma_map, label, id, gt, path):
    """Write the ROI table as ``df.to_csv(path, index=False)`` does (utils/t2map_utils.py:47-59).  ``path`` is the
    CSV file itself; building it from the BIDS layout stays with the caller (``get_img_path``, out of scope)."""
    tab = phantom_roi_table(t2_map, k_map, sigma_map, label, id, gt)
    cols = list(tab)
    with open(path, "w") as f:
        f.write(",".join(cols) + "\n")
        for i in range(len(gt)):
            row = []
            for c in cols:
                v = tab[c][i]
                row.append("" if isinstance(v, float) and np.isnan(v) else (repr(float(v)) if isinstance(v, (float, np.floating)) else str(v)))
            f.write(",".join(row) + "\n")
    return tab
