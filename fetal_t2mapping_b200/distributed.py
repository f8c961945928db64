"""Multi-GPU sharding of the fit: contiguous slabs of the masked-voxel list, one process per GPU.

Replaces the reference's only parallel runtime, ``multiprocessing.Pool(processes=20).map`` over
``mask_indices`` (run_t2mapping.py:442-443).  Voxels are independent, so there is no exchange step
during the fit; the only inter-GPU traffic is ONE final gather of the result vectors.  Slabs are cut
from the compacted masked list (balanced by masked count, 128-voxel aligned), not from z-slabs of the
volume, so a brain mask does not unbalance the GPUs.

* every rank needs ONLY ITS SLAB of the input: ``fit_slab_sharded`` takes the rows of the rank's own voxels
  (``slab_rows``), ``fit_voxels_sharded`` keeps the reference-shaped arguments (whole ``reshaped_t2w`` +
  ``mask_indices``) and reads only the slab's rows of them;
* the gather writes straight into the final layout: slabs are padded to one length L, every field is gathered with one
  ``all_gather_into_tensor`` into a ``[world * L]`` buffer whose first ``n_fit`` elements ARE the result (the padding sits
  behind the last slab); ``status`` travels as uint8;
* a rank whose fit raises (scipy's ``ValueError`` under ``--no_prior``, an index error, a CUDA error) tells the others
  before the collective: an error flag is all-reduced first and every rank raises, as ``pool.map`` aborts the whole map
  in the reference -- nobody is left waiting in a collective;
* ``fit_voxels_fused_gather``: no collective at all -- the fit kernels' epilogues store their slab straight into the
  root GPU's buffer over NVLink (CUDA IPC peer mapping);
* ``SlabPipeline``: a stream of sharded jobs with the (NCCL) gather of job i overlapping the fit of job i+1 (double-buffered);
* ``FusedAllGather``: the all-gather fused into the kernels -- every rank's epilogue stores its slab into EVERY rank's buffer
  over NVLink, one tiny all-reduce as the barrier (the fastest form measured: profiles/r02_notes.md).

Host logic only -- the fit of a slab is ``api.fit_voxels_batch`` (CUDA).  ``fit_fn`` is injectable so the partition /
gather / error logic is testable with ``gloo`` on CPU.
"""
from __future__ import annotations

import os

import numpy as np

__all__ = ["slab_bounds", "slab_length", "fit_voxels_sharded", "fit_slab_sharded", "gather_slabs", "gather_fields",
           "fit_voxels_fused_gather", "ShardError", "SlabPipeline", "FusedAllGather"]

ALIGN = 128
FIELDS = ("t2", "k", "sigma", "res", "status")


class ShardError(RuntimeError):
    """Another rank's fit failed; this rank's own slab was fine."""


def slab_length(n_fit: int, world: int, align: int = ALIGN) -> int:
    """The common (padded) slab length L: ceil(n_fit / world) rounded up to ``align``."""
    per = -(-max(n_fit, 0) // world)
    return max(align, -(-per // align) * align)


def slab_bounds(n_fit: int, world: int, align: int = ALIGN):
    """``world`` contiguous [start, stop) slabs of ``range(n_fit)``: slab r = [r L, (r+1) L) clipped to n_fit, L =
    ``slab_length``.  Every slab but the last non-empty one has exactly L voxels, so the concatenation of L-padded slabs is
    the full vector followed by padding (the all-gather then needs no stitching); sizes differ by less than L only at the
    tail; trailing slabs may be empty for tiny inputs."""
    if world < 1:
        raise ValueError("world must be >= 1")
    L = slab_length(n_fit, world, align)
    return [(min(n_fit, r * L), min(n_fit, (r + 1) * L)) for r in range(world)]


def _agree_or_raise(exc, device, group):
    """Exchange an error flag BEFORE any collective that depends on every rank having fitted its slab: all ranks raise if
    any rank failed (the reference's pool.map aborts the whole map when one voxel raises)."""
    import torch
    import torch.distributed as dist
    code = 0 if exc is None else (2 if isinstance(exc, ValueError) else 3 if isinstance(exc, IndexError) else 1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        flag = torch.tensor([code], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        worst = int(flag.item())
    else:
        worst = code
    if exc is not None:
        raise exc
    if worst:
        from .api import BOUNDS_ERROR
        if worst == 2:
            raise ValueError(BOUNDS_ERROR)          # what every rank of the reference's map would have seen
        if worst == 3:
            raise IndexError("mask_indices out of range (on another rank)")
        raise ShardError("the fit failed on another rank")


def gather_fields(local: dict, n_fit: int, group=None):
    """All-gather the slab vectors of ``local`` (name -> 1-D tensor of this rank's slab, any dtype) into full vectors on
    every rank.  One ``all_gather_into_tensor`` per field straight into the final buffer: slab r lands at [r L, (r+1) L),
    which is its place in the full vector; the result is the first ``n_fit`` elements -- no stitch copies."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    L = slab_length(n_fit, world)
    a, b = slab_bounds(n_fit, world)[rank]
    out = {}
    for name, v in local.items():
        assert v.dim() == 1 and v.shape[0] == b - a, "local slab does not match the partition"
        full = torch.empty(world * L, dtype=v.dtype, device=v.device)
        mine = full[rank * L:(rank + 1) * L]
        mine[:b - a].copy_(v)                      # the one copy: the slab into its place of the final buffer
        if b - a < L:
            mine[b - a:].zero_()
        # NCCL gathers in place (the input is this rank's chunk of the output); gloo (CPU tests) gets its own input buffer
        dist.all_gather_into_tensor(full, mine if v.is_cuda else mine.clone(), group=group)
        out[name] = full[:n_fit]
    return out


def gather_slabs(local, bounds, group=None):
    """All-gather the slabs of a 2-D tensor ``local`` [C, m_r] (one dtype) into [C, n_fit] on every rank (kept for callers
    that hold their fields stacked; ``gather_fields`` avoids the stacking)."""
    import torch
    n_fit = bounds[-1][1]
    got = gather_fields({str(c): local[c].contiguous() for c in range(local.shape[0])}, n_fit, group)
    return torch.stack([got[str(c)] for c in range(local.shape[0])])


def _as_tensor(x):
    import torch
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def _finish(r, exc, n_fit, bounds, rank, world, gather, group):
    import torch
    device = None
    if r is not None:
        device = _as_tensor(r.t2).device
    elif torch.cuda.is_available() and torch.distributed.is_initialized() and torch.distributed.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    _agree_or_raise(exc, device, group)
    local = {"t2": _as_tensor(r.t2).float(), "k": _as_tensor(r.k).float(), "sigma": _as_tensor(r.sigma).float(),
             "res": _as_tensor(r.res).float(), "status": _as_tensor(r.status).to(torch.uint8)}
    if world == 1 or not gather:
        out = dict(local) if world == 1 else {"local": local}
    else:
        out = gather_fields(local, n_fit, group)
    out.update(bounds=bounds, rank=rank)
    return out


def fit_slab_sharded(slab_rows, n_fit, TEeffs, fit, fit_params, prior=True, norm=False, *, group=None, fit_fn=None,
                     gather=True, **kw):
    """The product's multi-GPU call when every rank holds ONLY ITS OWN SLAB: ``slab_rows`` = the float32 rows ``[m_r, E]`` of
    the voxels ``slab_bounds(n_fit, world)[rank]`` of the masked list (what a loader hands each GPU).  Fits them and
    all-gathers (t2, k, sigma, res, status) into full-length vectors on every rank."""
    import torch.distributed as dist
    if fit_fn is None:
        from .api import fit_voxels_batch as fit_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bounds = slab_bounds(int(n_fit), world)
    a, b = bounds[rank]
    r, exc = None, None
    try:
        if slab_rows.shape[0] != b - a:
            raise IndexError(f"rank {rank} holds {slab_rows.shape[0]} rows, its slab has {b - a}")
        r = fit_fn(slab_rows, None, TEeffs, fit, fit_params, prior, norm, **kw)
    except Exception as e:                       # noqa: BLE001 -- exchanged with the other ranks, then re-raised
        exc = e
    return _finish(r, exc, int(n_fit), bounds, rank, world, gather, group)


def fit_voxels_sharded(reshaped_t2w, mask_indices, TEeffs, fit, fit_params, prior=True, norm=False, *, group=None,
                       fit_fn=None, gather=True, **kw):
    """The reference-shaped arguments on every rank (``reshaped_t2w`` may be a memory-mapped / host array: only the rows of
    the rank's own slab are read); rank r fits slab r of the masked list and the (t2, k, sigma, res, status) vectors are
    all-gathered.  Returns a dict of full-length arrays on every rank (``gather=False``: the local slab and its bounds).

    ``fit_fn(reshaped_t2w, mask_indices_slab, TEeffs, fit, fit_params, prior, norm, **kw)`` must return an
    object with ``t2, k, sigma, res, status`` (default: the CUDA path, ``api.fit_voxels_batch``).  If the fit raises on any
    rank, every rank raises (no rank is left in the collective)."""
    import torch.distributed as dist
    if fit_fn is None:
        from .api import fit_voxels_batch as fit_fn
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_fit = int(mask_indices.shape[0])
    bounds = slab_bounds(n_fit, world)
    a, b = bounds[rank]
    r, exc = None, None
    try:
        r = fit_fn(reshaped_t2w, mask_indices[a:b], TEeffs, fit, fit_params, prior, norm, **kw)
    except Exception as e:                       # noqa: BLE001 -- exchanged with the other ranks, then re-raised
        exc = e
    return _finish(r, exc, n_fit, bounds, rank, world, gather, group)


class _DeviceBytes:
    """Raw device memory as a ``__cuda_array_interface__`` provider (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def fit_voxels_fused_gather(reshaped_t2w, mask_indices, TEeffs, fit, fit_params, prior=True, norm=False, *, root=0, group=None,
                            solver="auto"):
    """The fit with the final gather FUSED into the kernels: rank r fits slab r of ``mask_indices`` and its kernel's
    epilogue stores (t2, k, sigma, res, status) straight into the ROOT GPU's full-length buffer over NVLink (peer stores
    through a CUDA-IPC mapping; no collective launch, the transfer overlaps the fit).  Device tensors in (each rank reads
    only its slab's rows).  Returns the full-length tensors on ``root`` (None on the other ranks) -- the reference's
    consumer of the maps is one process (NIfTI writers, run_t2mapping.py:471-479).  Raises on every rank what the reference
    raises (``ValueError`` for the --no_prior bounds, ``IndexError``) if any rank's slab holds such a voxel."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _abi
    from .api import _device_counts, _raise_for_counts, fit_voxels_into, init
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
        exc = None
        counts = _device_counts(torch, reshaped_t2w.device)
        try:
            if b > a:
                out = {"t2": base + 4 * a, "k": base + 4 * (n_fit + a), "sigma": base + 4 * (2 * n_fit + a),
                       "res": base + 4 * (3 * n_fit + a), "status": base + 16 * n_fit + a}
                fit_voxels_into(reshaped_t2w, mask_indices[a:b], TEeffs, fit, fit_params, prior, norm, out, solver=solver,
                                counts=counts)
            torch.cuda.synchronize()
            _raise_for_counts([int(v) for v in counts.cpu()])
        except Exception as e:                   # noqa: BLE001 -- exchanged with the other ranks, then re-raised
            exc = e
        _agree_or_raise(exc, reshaped_t2w.device, group)   # (an all-reduce: also the point where every rank's stores have landed)
        dist.barrier(group=group)
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


class SlabPipeline:
    """A STREAM of sharded jobs of one shape (e.g. the volumes of a series, each cut into ``world`` slabs): the all-gather of
    job i runs on NCCL's stream while the fit of job i+1 runs on the compute stream (``depth`` buffer sets, default 2).

    One job alone costs fit + gather; a stream of jobs costs max(fit, gather) per job.  Every rank RECEIVES (world - 1) slabs
    per job, so at 8 GPUs the gather of c2-sized slabs (7 x 21 MB per rank) is the longer of the two whatever the fit costs --
    overlapping is what is left to do about it.

    ``submit(rows, ...)`` enqueues the fit of this rank's slab (compact results straight into its chunk of the slot's gather
    buffers) and the asynchronous in-place all-gathers, and returns the slot; ``result(slot)`` makes the current stream wait
    for that slot's gathers and returns the full vectors (views of the slot's buffers: valid until the slot is reused,
    ``depth`` submits later).  Status histograms are accumulated per slot (``counts(slot)``)."""

    def __init__(self, n_fit, fit, *, depth=2, fields=None, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_fit, self.fit = int(n_fit), fit
        self.L = slab_length(self.n_fit, self.world)
        self.a, self.b = slab_bounds(self.n_fit, self.world)[self.rank]
        mono = fit == "gaussian"
        self.fields = tuple(fields) if fields else (("t2", "k", "res", "status") if mono else ("t2", "k", "sigma", "res", "status"))
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.slots = []
        for _ in range(max(1, depth)):
            bufs = {n: torch.zeros(self.world * self.L, dtype=torch.uint8 if n == "status" else torch.float32, device=dev)
                    for n in self.fields}
            mine = {n: bufs[n][self.rank * self.L:(self.rank + 1) * self.L] for n in self.fields}
            self.slots.append({"bufs": bufs, "mine": mine, "works": [], "counts": torch.zeros(4, dtype=torch.int64, device=dev)})
        self.next = 0

    def submit(self, slab_rows, TEeffs, fit_params, prior=True, norm=False, *, solver="auto"):
        from .api import fit_voxels_into
        s = self.next
        self.next = (self.next + 1) % len(self.slots)
        slot = self.slots[s]
        for w in slot["works"]:                     # the slot's previous gathers must have finished before the fit overwrites it
            w.wait()
        slot["works"] = []
        if slab_rows.shape[0] != self.b - self.a:
            raise IndexError(f"rank {self.rank} holds {slab_rows.shape[0]} rows, its slab has {self.b - self.a}")
        slot["counts"].zero_()
        out = {n: slot["mine"][n].data_ptr() for n in ("t2", "k", "res")}
        for n in ("sigma", "status"):
            if n in slot["mine"]:
                out[n] = slot["mine"][n].data_ptr()
        extra = {}
        if self.fit != "gaussian" and "sigma" not in slot["mine"]:
            extra["sigma"] = self.torch.empty(max(self.b - self.a, 1), dtype=self.torch.float32, device=slab_rows.device)
            out["sigma"] = extra["sigma"].data_ptr()
        if self.b > self.a:
            fit_voxels_into(slab_rows, None, TEeffs, self.fit, fit_params, prior, norm, out, solver=solver, counts=slot["counts"])
        slot["keep"] = extra
        # asynchronous collectives: NCCL's stream waits for the fit just enqueued; the compute stream does NOT wait for them
        slot["works"] = [self.dist.all_gather_into_tensor(slot["bufs"][n], slot["mine"][n], group=self.group, async_op=True)
                         for n in self.fields]
        return s

    def result(self, s, check=False):
        slot = self.slots[s]
        for w in slot["works"]:
            w.wait()                                # the current stream waits for the gathers (no host synchronisation)
        slot["works"] = []
        if check:                                   # synchronises: raise what the reference raises, on every rank
            exc = None
            try:
                from .api import _raise_for_counts
                _raise_for_counts([int(v) for v in slot["counts"].cpu()])
            except Exception as e:                  # noqa: BLE001
                exc = e
            _agree_or_raise(exc, slot["counts"].device, self.group)
        return {n: slot["bufs"][n][:self.n_fit] for n in self.fields}

    def drain(self):
        for s in range(len(self.slots)):
            for w in self.slots[s]["works"]:
                w.wait()
            self.slots[s]["works"] = []


class FusedAllGather:
    """Sharded jobs of one shape with the all-gather FUSED INTO THE FIT KERNELS: every rank's kernel epilogue stores its slab
    of (t2, k, [sigma,] res, status) into its own buffer AND, over NVLink, into the same slab of every peer's buffer (CUDA-IPC
    mappings, ``t2fit_outputs.dup_*``); a one-element NCCL all-reduce, stream-ordered after the kernels, is the cross-rank
    barrier that completes the gather.  No collective moves data: the transfer overlaps the fit, and a pass costs the fit plus
    ~15 us at 2 GPUs where NCCL's all-gather of the same 21 MB costs 77 us (profiles/r02_notes.md).  Every rank still
    RECEIVES (world - 1) slabs per job: at 8 GPUs the NVLink ingest (~150 MB per rank for c2-sized slabs) bounds the pass.

    ``multicast="on"``: the buffers live in torch symmetric memory and the float fields are stored ONCE per value to the
    buffers' NVSwitch multicast address (the switch replicates the store into every rank's buffer; a plain ``st.global`` is
    what ``multimem.st.weak`` assembles to on sm_100a).  Bit-identical, but measured slower than the unicast stores at 4 GPUs,
    hence off by default.

    ``depth`` buffer sets (default 2) let a consumer read job i while job i+1 is being written.  ``submit(rows, ...)`` ->
    slot; ``result(slot)`` -> full-length tensors (views of this rank's buffer, valid until the slot is reused)."""

    def __init__(self, n_fit, fit, *, depth=2, group=None, device=None, gather=None, multicast=None):
        """``gather``: the fields every rank receives from every peer (default: all of them).  ``("t2", "k")`` is the
        literal "final gather of the parameter maps" of BASELINE.json (T2 and S0): res / sigma / status of a slab then stay
        with its owner (``result()`` returns the owner's slab of those, zeros elsewhere) and 8 instead of 13-17 bytes per
        voxel cross NVLink."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _abi
        from .api import init
        self.torch, self.dist, self.group, self.C, self._abi = torch, dist, group, C, _abi
        self.gather = None if gather is None else tuple(gather)
        self.lib = init()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world - 1 > _abi.MAX_DUP:
            raise ValueError(f"the fused all-gather serves up to {_abi.MAX_DUP + 1} GPUs")
        self.n_fit, self.fit = int(n_fit), fit
        self.L = slab_length(self.n_fit, self.world)
        self.a, self.b = slab_bounds(self.n_fit, self.world)[self.rank]
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        tot = self.world * self.L
        self.fields = ("t2", "k", "res", "status") if fit == "gaussian" else ("t2", "k", "sigma", "res", "status")
        self.off, o = {}, 0
        for n in self.fields:
            self.off[n] = o
            o += tot * (1 if n == "status" else 4)
        self.nbytes = (o + 255) // 256 * 256
        self.sets = []
        # opt-in (multicast="on" or T2FIT_MULTICAST=on): measured SLOWER than the unicast peer stores on 4 x B200 (c2-sized slabs,
        # same box: 142.8 vs 122.6 us per pass, 442 vs 515 GB/s received per rank -- profiles/r02_notes.md section 11)
        want_mc = {"on": True, "1": True}.get(
            str(multicast if multicast is not None else os.environ.get("T2FIT_MULTICAST", "off")).lower(), False)
        self.multicast = False
        if want_mc:                                         # all slots or none: the two kinds of buffer sets are never mixed
            for _ in range(max(1, depth)):
                st = self._alloc_multicast(group)
                if st is None:
                    self.sets = []
                    break
                self.sets.append(st)
            self.multicast = bool(self.sets)
        while len(self.sets) < max(1, depth):
            self.sets.append(self._alloc_ipc(group))
        for st in self.sets:
            st["raw"].zero_()
            st.update(counts=torch.zeros(4, dtype=torch.int64, device=self.dev), call=None)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        torch.cuda.synchronize()
        dist.barrier(group=group)
        self.next = 0

    def _alloc_ipc(self, group):
        """One buffer set from the library's own allocator, mapped into every peer through CUDA-IPC (t2fit_shared_*)."""
        C, _abi, torch, dist = self.C, self._abi, self.torch, self.dist
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        _abi.check(self.lib, self.lib.t2fit_shared_alloc(self.nbytes, C.byref(ptr), handle), "t2fit_shared_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        peers = {}
        for r in range(self.world):
            if r != self.rank:
                pp = C.c_void_p()
                _abi.check(self.lib, self.lib.t2fit_shared_open(handles[r], C.byref(pp)), "t2fit_shared_open")
                peers[r] = pp
        raw = torch.as_tensor(_DeviceBytes(ptr.value, self.nbytes), device=self.dev)
        return {"ptr": ptr, "peers": peers, "raw": raw, "mc": 0, "symm": None}

    def _all_ok(self, ok, group):
        """True on every rank only if ``ok`` is true on every rank."""
        f = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(f, op=self.dist.ReduceOp.MIN, group=group)
        return bool(int(f.item()))

    def _alloc_multicast(self, group):
        """One buffer set in torch's symmetric memory (plumbing: CUDA VMM allocation, handle exchange, NVSwitch multicast
        binding) when the node's fabric offers a MULTICAST address for it: a store to that address is replicated by the switch
        into the same offset of every rank's buffer, so a slab leaves its GPU once instead of (world - 1) times.  Every step is
        agreed across the ranks; None (on all of them) = use the CUDA-IPC peer mappings.  The address is probed before use:
        every rank writes its rank number through it and all ranks must find all numbers in their own buffer."""
        torch, dist = self.torch, self.dist
        g = group if group is not None else dist.group.WORLD
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
        except Exception:                                   # noqa: BLE001
            t = None
        if not self._all_ok(t is not None, group):
            return None
        try:
            hdl = symm.rendezvous(t, g)
            mc = int(hdl.multicast_ptr or 0)
            ptrs = [int(q) for q in hdl.buffer_ptrs]
        except Exception:                                   # noqa: BLE001
            hdl, mc, ptrs = None, 0, []
        if not self._all_ok(mc != 0 and len(ptrs) == self.world and ptrs[self.rank] == t.data_ptr(), group):
            return None
        # probe: 256 bytes per rank at the start of the buffer
        t[:256 * self.world].zero_()
        torch.cuda.synchronize()
        dist.barrier(group=group)
        torch.as_tensor(_DeviceBytes(mc + 256 * self.rank, 256), device=self.dev).fill_(self.rank + 1)
        torch.cuda.synchronize()
        dist.barrier(group=group)
        seen = t[:256 * self.world].view(self.world, 256)
        want = torch.arange(1, self.world + 1, dtype=torch.uint8, device=self.dev)[:, None].expand(self.world, 256)
        ok = bool(torch.equal(seen, want))
        if not self._all_ok(ok, group):
            return None
        class _P:                                           # same attribute as the ctypes pointers of the IPC form
            def __init__(self, v):
                self.value = v
        return {"ptr": _P(ptrs[self.rank]), "peers": {r: _P(ptrs[r]) for r in range(self.world) if r != self.rank}, "raw": t,
                "mc": mc, "symm": hdl}

    def _build(self, st, slab_rows, TEeffs, fit_params, prior, norm, solver):
        """The t2fit_run arguments of a slot (built once per slot and input tensor; a pass then is two C calls)."""
        from .api import _fill_problem, resolve_solver
        _abi, C = self._abi, self.C
        p, o = _abi.Problem(), _abi.Outputs()
        keep = _fill_problem(p, self.fit, fit_params, TEeffs, prior, norm, 0, 0.0, "auto", resolve_solver(self.fit, solver))
        p.echoes, p.memory, p.layout, p.mask_idx = slab_rows.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS, None
        p.n_vox = p.n_fit = slab_rows.shape[0]

        def at(base, n):
            return base + self.off[n] + (1 if n == "status" else 4) * self.rank * self.L
        base = st["ptr"].value
        o.t2, o.k, o.res, o.status = at(base, "t2"), at(base, "k"), at(base, "res"), at(base, "status")
        o.sigma = at(base, "sigma") if "sigma" in self.off else None
        o.dense, o.counts_dev = 0, st["counts"].data_ptr()
        o.n_dup = len(st["peers"])
        sel = self.fields if self.gather is None else self.gather
        for j, r in enumerate(sorted(st["peers"])):
            pb = st["peers"][r].value
            for n, arr in (("t2", o.dup_t2), ("k", o.dup_k), ("res", o.dup_res), ("status", o.dup_status), ("sigma", o.dup_sigma)):
                arr[j] = at(pb, n) if (n in self.off and n in sel) else None
                if st["mc"] and n != "status":
                    # float fields through the multicast address: ONE store per value (destination 0), replicated by the
                    # switch into every rank's buffer; status (bytes; multimem stores are >= 32 bits wide) stays unicast
                    arr[j] = at(st["mc"], n) if (j == 0 and n in self.off and n in sel) else None
        return {"p": p, "o": o, "keep": (keep, slab_rows), "key": (slab_rows.data_ptr(), slab_rows.shape[0], id(fit_params), prior, norm, solver)}

    def submit(self, slab_rows, TEeffs, fit_params, prior=True, norm=False, *, solver="auto"):
        s = self.next
        self.next = (self.next + 1) % len(self.sets)
        st = self.sets[s]
        if slab_rows.shape[0] != self.b - self.a:
            raise IndexError(f"rank {self.rank} holds {slab_rows.shape[0]} rows, its slab has {self.b - self.a}")
        key = (slab_rows.data_ptr(), slab_rows.shape[0], id(fit_params), prior, norm, solver)
        if st["call"] is None or st["call"]["key"] != key:
            st["call"] = self._build(st, slab_rows, TEeffs, fit_params, prior, norm, solver)
        st["counts"].zero_()
        stream = self.torch.cuda.current_stream(self.dev).cuda_stream
        if self.b > self.a:
            rc = self.lib.t2fit_run(self.C.byref(st["call"]["p"]), self.C.byref(st["call"]["o"]), stream)
            if rc:
                raise self._abi.T2FitError((self.lib.t2fit_last_error() or b"").decode())
        # cross-rank barrier ordered after the kernels: when it completes here, every rank's kernel (and its peer stores) has
        self.dist.all_reduce(self.flag, group=self.group)
        return s

    def result(self, s, check=False):
        st = self.sets[s]
        if check:
            exc = None
            try:
                from .api import _raise_for_counts
                _raise_for_counts([int(v) for v in st["counts"].cpu()])
            except Exception as e:                  # noqa: BLE001
                exc = e
            _agree_or_raise(exc, self.dev, self.group)
        tot = self.world * self.L
        out = {}
        for n in self.fields:
            if n == "status":
                out[n] = st["raw"][self.off[n]:self.off[n] + tot][:self.n_fit]
            else:
                out[n] = st["raw"][self.off[n]:self.off[n] + 4 * tot].view(self.torch.float32)[:self.n_fit]
        return out

    def close(self):
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)
        for st in self.sets:
            if not st["mc"]:
                for pp in st["peers"].values():
                    self.lib.t2fit_shared_close(pp)
            st["raw"] = None
        self.dist.barrier(group=self.group)
        for st in self.sets:
            if not st["mc"]:
                self.lib.t2fit_shared_free(st["ptr"])
            st["symm"] = None
        self.sets = []
