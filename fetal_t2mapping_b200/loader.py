"""Batched, overlapped loader for the T2 fit: the block of ``process_t2maps`` that feeds the hot path.

Reference (run_t2mapping.py:365-400): per subject/session, read one recon volume and one mask volume per
echo time, ``np.stack`` them, union the masks (``np.sum(mask, axis=3) > 0``), optionally mask by the
phantom label (``--in_vitro_fast``: ``mask[label == 0] = 0``), flatten and cast to float32 (:411-412).

Here the per-TE volumes go to the GPU *as they come off disk* -- plane e of an ``[E, N]`` device buffer
(``T2FIT_LAYOUT_PLANES``), no interleaving ``np.stack`` on the host -- through page-locked staging
buffers on a copy stream, while the previous volume is being fitted on the compute stream:

    host cast into pinned planes -> H2D (copy stream) -> mask union + label masking -> ordered mask
    indices -> fit + residuals + zero-filled dense maps (one t2fit_run) -> D2H of the four maps

``depth`` (>= 2, default 3) staging slots are in flight.  Everything on the device runs through the C ABI
(``t2fit_mask_union``, ``t2fit_mask_indices``, ``t2fit_run``); torch provides pinned memory, streams
and events only.

``route="auto"`` (default) does the mask part on the HOST instead (``t2fit_host_mask_union_indices``: union, label
masking and ``np.where`` on the library's worker threads), and -- when the union covers at most half of the volume, a
brain mask covers ~10 % -- gathers only the masked voxels of every per-TE volume into a page-locked ``[E, M]`` block
(``t2fit_host_gather_planes``, ``T2FIT_LAYOUT_SOA``): the H2D copy shrinks from E volumes + E masks to the masked voxels,
one mask plane and the index vector, and the device part needs no host round trip (the voxel count is already known).
Denser masks keep the full per-TE planes.  ``route="device"`` is the all-device form described above.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from .api import BOUNDS_ERROR, _fill_problem, _run, _state, init, resolve_solver

__all__ = ["t2map_series", "VolumeMaps"]


@dataclass
class VolumeMaps:
    """What ``process_t2maps`` holds for one subject/session after the fit (:471-474): the four float32 maps
    ``[z,y,x]`` (zeros off-mask), plus the union mask and the number of fitted voxels."""
    t2: np.ndarray
    k: np.ndarray
    sigma: np.ndarray
    res: np.ndarray
    mask: np.ndarray
    n_fit: int
    failed: int


class _Slot:
    def __init__(self, torch, dev, n_echo, n_vox, mask_dtype, label_dtype):
        self.n_echo, self.n_vox = n_echo, n_vox
        self.h_planes = torch.empty((n_echo, n_vox), dtype=torch.float32, pin_memory=True)
        self.d_planes = torch.empty((n_echo, n_vox), dtype=torch.float32, device=dev)
        device_route = mask_dtype is not None              # route="device": per-TE masks (and the label) go to the GPU
        self.h_masks = torch.empty((n_echo, n_vox), dtype=mask_dtype, pin_memory=True) if device_route else None
        self.d_masks = torch.empty((n_echo, n_vox), dtype=mask_dtype, device=dev) if device_route else None
        self.h_label = torch.empty(n_vox, dtype=label_dtype, pin_memory=True) if label_dtype is not None else None
        self.d_label = torch.empty(n_vox, dtype=label_dtype, device=dev) if label_dtype is not None else None
        self.h_idx = None if device_route else torch.empty(n_vox, dtype=torch.int64, pin_memory=True)
        self.compact = False                               # echoes of the volume in flight are [E, n_fit] (SOA), not planes
        self.d_mask = torch.empty(n_vox, dtype=torch.uint8, device=dev)
        self.d_idx = torch.empty(n_vox, dtype=torch.int64, device=dev)
        self.d_maps = torch.empty((4, n_vox), dtype=torch.float32, device=dev)
        self.h_maps = None         # page-locked result block of the volume in flight; handed to the caller, not reused
        self.h_mask = None
        self.d_status = torch.empty(n_vox, dtype=torch.uint8, device=dev)
        self.h_cnt = torch.zeros(4, dtype=torch.int64, pin_memory=True)
        self.d_cnt = torch.zeros(4, dtype=torch.int64, device=dev)      # this volume's own status histogram (counts_dev)
        self.copied = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.shape3 = None
        self.has_label = False
        self.n_fit = 0
        self.busy = False


def _stage(torch, dst, a):
    """Cast / copy one volume into its page-locked plane (torch's copy is multi-threaded for large arrays)."""
    a = np.ascontiguousarray(a).reshape(-1)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    try:
        dst.copy_(torch.from_numpy(a))
    except (TypeError, RuntimeError):                      # dtypes torch cannot wrap (uint16, ...)
        np.copyto(dst.numpy(), a, casting="unsafe")


def _torch_dtype(torch, np_dtype):
    name = np.dtype(np_dtype).name
    if name == "bool":
        return torch.uint8
    if name in ("uint8", "int16", "int32", "float32", "float64"):
        return getattr(torch, name)
    if name == "uint16":
        return torch.int32        # widened on the host (torch has no pinned uint16 everywhere)
    return torch.float32


def _host_array(a, codes):
    """C-contiguous host array of a dtype the library's host functions read (others are widened as numpy would sum / cast them)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    if a.dtype.name not in codes:
        a = a.astype(np.float64 if a.dtype.kind in "iu" else np.float32)
    return a


def t2map_series(volumes, TEeffs, fit, fit_params, prior=True, norm=False, *, fast=False, solver="auto", depth=3, route="auto"):
    """Fit a series of subjects/sessions, overlapping each volume's staging with the previous volume's fit.

    ``volumes``: iterable of ``(t2w_list, mask_list)`` or ``(t2w_list, mask_list, label)`` -- per-TE recon and
    mask arrays ``[z,y,x]`` in ascending echo-time order as ``process_t2maps`` reads them (:365-381), all of one
    shape per volume; ``label`` the phantom label image or None.  ``fast`` = ``--in_vitro_fast`` (mask by label,
    :393-400).  Yields one :class:`VolumeMaps` per volume, in order.  Raises ``ValueError`` where the reference's
    map would abort (scipy bounds error under ``--no_prior``).  ``route``: "auto" (mask part on the host, masked voxels only
    over PCIe for sparse masks) or "device" (whole volumes and masks to the GPU) -- module docstring."""
    import torch
    if route not in ("auto", "device"):
        raise ValueError(f"unknown route {route!r}")
    lib = init()
    dev = torch.device("cuda", _state["device"])
    solver = resolve_solver(fit, solver)
    te = np.asarray(TEeffs, np.float64).reshape(-1)
    n_echo = te.size
    copy_stream = torch.cuda.Stream(dev)
    compute = torch.cuda.Stream(dev)
    slots = {}
    pending = []                   # slots whose results have not been yielded yet, in order

    def finish(s):
        s.done.synchronize()
        s.busy = False
        _, nonfinite, gave_up, bad_bounds = (int(v) for v in s.h_cnt)
        if bad_bounds > 0:
            raise ValueError(BOUNDS_ERROR)
        maps, mk = s.h_maps.numpy(), s.h_mask.numpy()          # views of the page-locked blocks: no copy; the blocks go
        s.h_maps = s.h_mask = None                               # back to torch's pinned pool when the caller drops them
        return VolumeMaps(*(maps[i].reshape(s.shape3) for i in range(4)),
                          mask=mk.reshape(s.shape3).view(np.bool_), n_fit=s.n_fit, failed=nonfinite + gave_up)

    def launch(s):
        """Device part of a staged volume: mask union (+ label), ordered indices, fit into zero-filled dense maps, D2H."""
        n_vox = s.n_vox
        compute.wait_event(s.copied)
        cs = compute.cuda_stream
        if s.d_masks is not None:                          # route="device": union, label masking and np.where on the GPU
            planes = (C.c_void_p * n_echo)(*[s.d_masks[e].data_ptr() for e in range(n_echo)])
            dt_code = _abi.DTYPES[str(s.d_masks.dtype).replace("torch.", "")]
            l_code = _abi.DTYPES[str(s.d_label.dtype).replace("torch.", "")] if s.has_label else 0
            _abi.check(lib, lib.t2fit_mask_union(planes, n_echo, dt_code, s.d_label.data_ptr() if s.has_label else None, l_code,
                                                 n_vox, s.d_mask.data_ptr(), cs), "t2fit_mask_union")
            n = C.c_int64()
            # returns the count, i.e. waits for this volume's H2D -- which has had the staging time of the NEXT volume to finish
            _abi.check(lib, lib.t2fit_mask_indices(s.d_mask.data_ptr(), n_vox, 1, s.d_idx.data_ptr(), C.byref(n), cs),
                       "t2fit_mask_indices")
            s.n_fit = int(n.value)
        p, o = _abi.Problem(), _abi.Outputs()
        keep = _fill_problem(p, fit, fit_params, te, prior, norm, 0, 0.0, "auto", solver)
        p.echoes, p.memory = s.d_planes.data_ptr(), _abi.MEM_DEVICE
        p.layout, p.ld = (_abi.LAYOUT_SOA, max(1, s.n_fit)) if s.compact else (_abi.LAYOUT_PLANES, n_vox)
        p.mask_idx, p.n_vox, p.n_fit = s.d_idx.data_ptr(), n_vox, s.n_fit
        o.t2, o.k, o.sigma, o.res = (s.d_maps[i].data_ptr() for i in range(4))
        o.dense, o.zero_fill_mask = 1, s.d_mask.data_ptr()
        o.status = s.d_status.data_ptr()
        with torch.cuda.stream(compute):
            s.d_cnt.zero_()
        o.counts_dev = s.d_cnt.data_ptr()
        _run(lib, p, o, cs)
        del keep
        with torch.cuda.stream(compute):
            s.h_cnt.copy_(s.d_cnt, non_blocking=True)
            s.h_maps.copy_(s.d_maps, non_blocking=True)
            if s.d_masks is not None:
                s.h_mask.copy_(s.d_mask, non_blocking=True)    # (the host routes computed the union into h_mask themselves)
            s.done.record(compute)
        pending.append(s)

    # Software pipeline over the volumes: volume v is staged (host cast into page-locked planes, H2D enqueued on the copy
    # stream) BEFORE the device part of volume v-1 is launched, so the one host wait of the device part (the voxel count of
    # t2fit_mask_indices) meets an H2D that finished while v was being staged, and v's H2D overlaps the fit and D2H of v-1.
    staged = None                  # the slot whose H2D is enqueued but whose device part has not been launched yet
    for item in volumes:
        t2w_list, mask_list = item[0], item[1]
        label = item[2] if len(item) > 2 else None
        if len(t2w_list) != n_echo or len(mask_list) != n_echo:
            raise ValueError(f"{n_echo} echo times but {len(t2w_list)} volumes / {len(mask_list)} masks")
        shape3 = tuple(np.shape(t2w_list[0]))
        n_vox = int(np.prod(shape3))
        on_host = route == "auto"
        m_dt = None if on_host else _torch_dtype(torch, np.asarray(mask_list[0]).dtype)
        l_dt = _torch_dtype(torch, np.asarray(label).dtype) if (label is not None and fast and not on_host) else None
        key = (n_vox, m_dt, l_dt)
        ring = slots.setdefault(key, [])
        s = next((x for x in ring if not x.busy), None)
        if s is None and len(ring) < max(2, depth):
            s = _Slot(torch, dev, n_echo, n_vox, m_dt, l_dt)
            ring.append(s)
        while s is None:                                   # all slots of this shape in flight: drain the oldest
            if not pending:
                launch(staged)
                staged = None
            yield finish(pending.pop(0))
            s = next((x for x in ring if not x.busy), None)
        s.busy, s.shape3 = True, shape3
        for e in range(n_echo):
            if np.shape(t2w_list[e]) != shape3 or np.shape(mask_list[e]) != shape3:
                raise ValueError("all per-TE volumes and masks of one subject must have the same shape")
        s.h_maps = torch.empty((4, n_vox), dtype=torch.float32, pin_memory=True)
        s.h_mask = torch.empty(n_vox, dtype=torch.uint8, pin_memory=True)
        if on_host:
            # ---- host: union + label masking + np.where; then either the masked voxels only ([E, M], cast to float32) or,
            # for a dense mask, the whole planes; H2D on the copy stream; the device part follows at once (no count to wait for)
            masks = [_host_array(mask_list[e], _abi.DTYPES) for e in range(n_echo)]
            if len({m.dtype for m in masks}) > 1:
                masks = [m.astype(np.float64) for m in masks]
            lab = _host_array(label, _abi.DTYPES) if (label is not None and fast) else None
            if lab is not None and lab.shape != shape3:
                raise ValueError("the label image must have the shape of the volumes")
            n = C.c_int64()
            _abi.check(lib, lib.t2fit_host_mask_union_indices(
                (C.c_void_p * n_echo)(*[m.ctypes.data for m in masks]), n_echo, _abi.DTYPES[masks[0].dtype.name],
                lab.ctypes.data if lab is not None else None, _abi.DTYPES[lab.dtype.name] if lab is not None else 0,
                n_vox, s.h_mask.data_ptr(), s.h_idx.data_ptr(), C.byref(n)), "t2fit_host_mask_union_indices")
            s.has_label = False
            s.n_fit = m_fit = int(n.value)
            s.compact = 2 * m_fit <= n_vox
            if s.compact:
                vols = [_host_array(t2w_list[e], _abi.DTYPES) for e in range(n_echo)]
                if len({v.dtype for v in vols}) > 1:
                    vols = [v.astype(np.float32) for v in vols]
                _abi.check(lib, lib.t2fit_host_gather_planes(
                    (C.c_void_p * n_echo)(*[v.ctypes.data for v in vols]), n_echo, _abi.DTYPES[vols[0].dtype.name],
                    s.h_idx.data_ptr(), m_fit, n_vox, s.h_planes.data_ptr(), max(1, m_fit)), "t2fit_host_gather_planes")
            else:
                for e in range(n_echo):
                    _stage(torch, s.h_planes[e], np.asarray(t2w_list[e]))
            n_copy = n_echo * m_fit if s.compact else n_echo * n_vox
            with torch.cuda.stream(copy_stream):
                s.d_planes.view(-1)[:n_copy].copy_(s.h_planes.view(-1)[:n_copy], non_blocking=True)
                s.d_mask.copy_(s.h_mask, non_blocking=True)
                s.d_idx[:m_fit].copy_(s.h_idx[:m_fit], non_blocking=True)
                s.copied.record(copy_stream)
            launch(s)
        else:
            # ---- host: cast into page-locked planes (the .astype(np.float32) of :411), enqueue H2D on the copy stream
            s.compact = False
            for e in range(n_echo):
                _stage(torch, s.h_planes[e], np.asarray(t2w_list[e]))
                _stage(torch, s.h_masks[e], np.asarray(mask_list[e]))
            s.has_label = l_dt is not None
            if s.has_label:
                _stage(torch, s.h_label, np.asarray(label))
            with torch.cuda.stream(copy_stream):
                s.d_planes.copy_(s.h_planes, non_blocking=True)
                s.d_masks.copy_(s.h_masks, non_blocking=True)
                if s.has_label:
                    s.d_label.copy_(s.h_label, non_blocking=True)
                s.copied.record(copy_stream)
            # ---- device part of the PREVIOUS volume
            if staged is not None:
                launch(staged)
            staged = s
        while len(pending) >= max(2, depth):
            yield finish(pending.pop(0))
    if staged is not None:
        launch(staged)
    while pending:
        yield finish(pending.pop(0))
