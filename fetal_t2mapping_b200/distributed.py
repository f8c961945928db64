"""Multi-GPU sharding of the fit: contiguous slabs of the masked-voxel list, one process per GPU.

Replaces the reference's only parallel runtime, ``multiprocessing.Pool(processes=20).map`` over
``mask_indices`` (run_t2mapping.py:442-443).  Voxels are independent, so there is no exchange step
during the fit; the only inter-GPU traffic is ONE final gather of the parameter vectors
(NCCL over NVLink, ``all_gather_into_tensor``).  Slabs are cut from the compacted masked list
(balanced by masked count, 128-voxel aligned), not from z-slabs of the volume, so a brain mask
does not unbalance the GPUs.

Host logic only -- the fit of a slab is ``api.fit_voxels_batch`` (CUDA).  ``fit_fn`` is injectable
so the partition / gather logic is testable with ``gloo`` on CPU.
"""
from __future__ import annotations

import numpy as np

__all__ = ["slab_bounds", "fit_voxels_sharded", "gather_slabs", "fit_voxels_fused_gather"]

ALIGN = 128


def slab_bounds(n_fit: int, world: int, align: int = ALIGN):
    """``world`` contiguous [start, stop) slabs of ``range(n_fit)``; interior cuts are multiples of
    ``align``; sizes differ by at most ``align``; trailing slabs may be empty for tiny inputs."""
    if world < 1:
        raise ValueError("world must be >= 1")
    blocks = -(-n_fit // align)
    cuts = [min(n_fit, ((blocks * r) // world) * align) for r in range(world + 1)]
    cuts[-1] = n_fit
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_slabs(local, bounds, group=None):
    """All-gather equally padded slabs of a float32 tensor ``local`` [C, m_r] and stitch them into
    [C, n_fit] on every rank.  One collective (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [b - a for a, b in bounds]
    mx = max(max(sizes), 1)
    c = local.shape[0]
    assert local.shape[1] == sizes[rank], "local slab does not match the partition"
    pad = torch.zeros((c, mx), dtype=local.dtype, device=local.device)
    pad[:, :sizes[rank]] = local
    flat = torch.empty((world * c, mx), dtype=local.dtype, device=local.device)   # concatenation along dim 0
    dist.all_gather_into_tensor(flat, pad, group=group)
    out = flat.view(world, c, mx)
    full = torch.empty((c, bounds[-1][1]), dtype=local.dtype, device=local.device)
    for r, (a, b) in enumerate(bounds):
        full[:, a:b] = out[r, :, :b - a]
    return full


def fit_voxels_sharded(reshaped_t2w, mask_indices, TEeffs, fit, fit_params, prior=True, norm=False, *, group=None,
                       fit_fn=None, gather=True, **kw):
    """Every rank holds (or can read) ``reshaped_t2w`` / ``mask_indices``; rank r fits slab r of the
    masked list and the (t2, k, sigma, res, status) vectors are all-gathered.  Returns a dict of
    full-length arrays on every rank (``gather=False``: the local slab and its bounds only).

    ``fit_fn(reshaped_t2w, mask_indices_slab, TEeffs, fit, fit_params, prior, norm, **kw)`` must return an
    object with ``t2, k, sigma, res, status`` (default: the CUDA path, ``api.fit_voxels_batch``).
    """
    import torch
    import torch.distributed as dist
    if fit_fn is None:
        from .api import fit_voxels_batch as fit_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_fit = int(mask_indices.shape[0])
    bounds = slab_bounds(n_fit, world)
    a, b = bounds[rank]
    r = fit_fn(reshaped_t2w, mask_indices[a:b], TEeffs, fit, fit_params, prior, norm, **kw)

    def as_tensor(x):
        return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    local = torch.stack([as_tensor(r.t2).float(), as_tensor(r.k).float(), as_tensor(r.sigma).float(),
                         as_tensor(r.res).float(), as_tensor(r.status).float()])
    if not gather or world == 1:
        full = local
        if world > 1:
            return {"local": local, "bounds": bounds, "rank": rank}
    else:
        full = gather_slabs(local, bounds, group)
    return {"t2": full[0], "k": full[1], "sigma": full[2], "res": full[3], "status": full[4].to(torch.uint8),
            "bounds": bounds, "rank": rank}


class _DeviceBytes:
    """Raw device memory as a ``__cuda_array_interface__`` provider (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def fit_voxels_fused_gather(reshaped_t2w, mask_indices, TEeffs, fit, fit_params, prior=True, norm=False, *, root=0, group=None,
                            solver="auto"):
    """The fit with the final gather FUSED into the kernels: rank r fits slab r of ``mask_indices`` and its kernel's
    epilogue stores (t2, k, sigma, res, status) straight into the ROOT GPU's full-length buffer over NVLink (peer stores
    through a CUDA-IPC mapping; no collective launch, the transfer overlaps the fit).  Device tensors in.  Returns the
    full-length tensors on ``root`` (None on the other ranks) -- the reference's consumer of the maps is one process
    (NIfTI writers, run_t2mapping.py:471-479)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _abi
    from .api import fit_voxels_into, init
    lib = init()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_fit = int(mask_indices.shape[0])
    bounds = slab_bounds(n_fit, world)
    a, b = bounds[rank]
    nbytes = 4 * 4 * n_fit + n_fit                         # [t2 | k | sigma | res] float32, then status uint8
    ptr = C.c_void_p()
    payload = [None]
    if rank == root:
        handle = C.create_string_buffer(64)
        _abi.check(lib, lib.t2fit_shared_alloc(max(nbytes, 1), C.byref(ptr), handle), "t2fit_shared_alloc")
        payload = [handle.raw]
    dist.broadcast_object_list(payload, src=root, group=group)
    if rank != root:
        _abi.check(lib, lib.t2fit_shared_open(payload[0], C.byref(ptr)), "t2fit_shared_open")
    base = ptr.value
    try:
        if b > a:
            out = {"t2": base + 4 * a, "k": base + 4 * (n_fit + a), "sigma": base + 4 * (2 * n_fit + a),
                   "res": base + 4 * (3 * n_fit + a), "status": base + 16 * n_fit + a}
            fit_voxels_into(reshaped_t2w, mask_indices[a:b], TEeffs, fit, fit_params, prior, norm, out, solver=solver)
        torch.cuda.synchronize()
        dist.barrier(group=group)                          # every rank's stores have landed in the root's memory
        result = None
        if rank == root:
            raw = torch.as_tensor(_DeviceBytes(base, max(nbytes, 1)), device=reshaped_t2w.device)
            f = raw[:16 * n_fit].view(torch.float32).view(4, n_fit).clone()
            if fit == "gaussian":
                f[2].zero_()                               # the 2-parameter fit never writes sigma
            result = {"t2": f[0], "k": f[1], "sigma": f[2], "res": f[3], "status": raw[16 * n_fit:16 * n_fit + n_fit].clone(),
                      "bounds": bounds, "rank": rank}
            del raw
            torch.cuda.synchronize()
        dist.barrier(group=group)
        return result
    finally:
        if rank == root:
            lib.t2fit_shared_free(ptr)
        else:
            lib.t2fit_shared_close(ptr)
