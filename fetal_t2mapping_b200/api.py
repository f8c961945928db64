"""Host-side mirror of the reference's fit block, over the libt2fit C ABI.

Drop-in for the call pair inside ``process_t2maps`` (run_t2mapping.py:430-461)::

    all_results = pool.map(partial(fit_voxel, fit=, fit_params=, TEeffs=, reshaped_t2w=, prior=, norm=), mask_indices)
    res_map     = compute_residuals(reshaped_t2w, TEeffs, fit, norm, k_map, t2_map, sigma_map, res_map, mask_indices, mask)

``fit_voxels_batch`` takes exactly those arguments (same names, same meaning) and returns the
same information for every masked voxel at once; ``t2map_volume`` wraps the whole hot block
(mask union, flatten, fit, scatter, residual map; :383-386, :411-461, :471-473).

numpy arrays are treated as host memory (the library stages them through pinned buffers);
torch CUDA tensors are treated as device memory (zero-copy, asynchronous on the current torch
stream).  The fit itself only exists as sm_100a CUDA kernels: without the built library or
without a GPU these functions raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _abi
from .presets import NO_PRIOR_K_UB, NO_PRIOR_T2_BOUNDS

__all__ = ["fit_voxels_batch", "fit_voxels_into", "check_counts", "t2map_volume", "compute_residuals", "FitResult", "init", "shutdown", "device_info",
           "mask_indices_device", "work_model", "pinned_array", "BOUNDS_ERROR"]

BOUNDS_ERROR = "An upper bound is less than the corresponding lower bound."   # scipy's text

_state = {"device": None}


def init(device: int | None = None):
    """Bind this process to one GPU (default: LOCAL_RANK, else torch's current device, else 0)."""
    lib = _abi.load_library()
    if device is None:
        if _state["device"] is not None:
            return lib
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _state["device"] == device:
        return lib
    # one process per GPU: share the host cores between the ranks of this node for the staging threads of the host path
    lws = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    if lws > 1 and "T2FIT_HOST_THREADS" not in os.environ:
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        os.environ["T2FIT_HOST_THREADS"] = str(max(2, min(16, ncpu // lws)))
    _abi.check(lib, lib.t2fit_init(int(device)), "t2fit_init")
    _state["device"] = int(device)
    return lib


def shutdown():
    if _state["device"] is not None:
        _abi.load_library().t2fit_shutdown()
        _state["device"] = None


def device_info():
    lib = init()
    name = C.create_string_buffer(256)
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    _abi.check(lib, lib.t2fit_device_info(name, 256, C.byref(sm), C.byref(maj), C.byref(mnr)), "t2fit_device_info")
    return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value), "device": _state["device"]}


def work_model(fit: str, n_echo: int):
    """Algorithmic work of the shipped kernel per voxel (roofline accounting)."""
    lib = _abi.load_library()
    v = [C.c_double() for _ in range(5)]
    _abi.check(lib, lib.t2fit_work_model(_abi.MODELS[fit], n_echo, *[C.byref(x) for x in v]), "t2fit_work_model")
    keys = ("flop_per_pass", "mufu_per_pass", "flop_fixed", "mufu_fixed", "bytes_per_voxel")
    return {k: x.value for k, x in zip(keys, v)}


def _host_echoes(y, p):
    """The host echo array as the library takes it.  The reference casts the whole volume (`.astype(np.float32)`,
    run_t2mapping.py:411); a C-contiguous int16 / uint16 / int32 / float64 array -- what the NIfTI reader returns -- is
    passed as it is instead and cast by the library while it gathers the masked rows (t2fit_problem.echo_dtype)."""
    code = _abi.ECHO_DTYPES.get(y.dtype.name)
    if code is None or not y.flags.c_contiguous:
        y = np.ascontiguousarray(y, dtype=np.float32)
        code = 0
    p.echo_dtype = code
    return y


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def pinned_array(shape, dtype=np.float32, like=None):
    """A numpy array in page-locked host memory (backed by a torch pinned tensor that lives as long as the array).
    ``like``: copy this array into it.  When ``reshaped_t2w`` (and optionally ``mask_indices``) handed to
    :func:`fit_voxels_batch` are page-locked, the call needs no staging: the kernel gathers the masked rows straight
    from host memory over PCIe and stores the results straight back (``run_host_mapped`` in csrc/t2fit_kernels.cu)."""
    import torch
    if like is not None:
        like = np.asarray(like)
        shape, dtype = like.shape, like.dtype
    t = torch.empty(tuple(np.atleast_1d(shape)) if not isinstance(shape, tuple) else shape,
                    dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
    a = t.numpy()
    if like is not None:
        np.copyto(a, like)
    return a


def _host_block(m, fields):
    """All result arrays of one call as views into ONE page-locked block (one pinned allocation / DMA mapping instead
    of one per field).  ``fields``: list of (name, dtype).  Falls back to plain numpy without torch."""
    offs, total = {}, 0
    for name, dt in fields:
        total = (total + 255) // 256 * 256                       # 256-byte aligned views (vector stores, cache lines)
        offs[name] = total
        total += m * np.dtype(dt).itemsize
    try:
        import torch
        base = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True).numpy()
    except Exception:
        base = np.empty(max(total, 1), np.uint8)
    return {name: base[offs[name]:offs[name] + m * np.dtype(dt).itemsize].view(dt) for name, dt in fields}


@dataclass
class FitResult:
    """Everything ``pool.map(fit_voxel)`` + ``compute_residuals`` produce, for all masked voxels.

    ``results``            f64[M,P] in the reference order (k, T2[, sigma])       (:449)
    ``convergence_flags``  bool[M]   == ``result.success``                        (:450)
    ``num_iterations``     int32[M]  passes of the CUDA solver (not L-BFGS-B's)   (:451)
    ``final_errors``       f64[M]    mean squared error at the solution           (:452)
    ``res``                f32[M]    signed mean residual (compute_residuals)     (:461)
    ``status``             uint8[M]  0 ok / 1 non-finite input / 2 iteration cap / 3 bad bounds
    ``iteration_infos``    the callback traces (:180-234) when the call asked for them (``trace_cap``,
                           L-BFGS-B solver), else ``[]`` per voxel
    For device calls the float fields are torch CUDA tensors (float32).
    """
    t2: object
    k: object
    _sigma: object
    res: object
    fun: object
    nit: object
    status: object
    fit: str
    status_count: tuple = (0, 0, 0, 0)
    solver: str = "fast"
    trace_f: object = None       # float32[M, trace_cap]  f_val per L-BFGS-B iteration
    trace_step: object = None    # float32[M, trace_cap]  step_size (NaN for the first iteration)
    trace_len: object = None     # int32[M]               entries used = min(nit, trace_cap)

    @property
    def sigma(self):
        """sigma_map values; all zeros for the 2-parameter fit (run_t2mapping.py:417,457-458), created on first use."""
        if self._sigma is None:
            self._sigma = np.zeros(self.t2.shape[0], np.float32)
        return self._sigma

    @property
    def iteration_infos(self):
        """``iteration_info`` of every voxel as the reference's callbacks build it (run_t2mapping.py:180-234):
        a list of ``{'f_val', 'grad_norm': None, 'step_size'}`` per L-BFGS-B iteration."""
        m = int(self.nit.shape[0]) if self.nit is not None else int(self.t2.shape[0])
        if self.trace_len is None:
            return [[] for _ in range(m)]
        tn = np.asarray(self.trace_len.cpu() if _is_torch(self.trace_len) else self.trace_len)
        tf = np.asarray(self.trace_f.cpu() if _is_torch(self.trace_f) else self.trace_f)
        ts = np.asarray(self.trace_step.cpu() if _is_torch(self.trace_step) else self.trace_step)
        return [[{"f_val": float(tf[i, j]), "grad_norm": None, "step_size": float(ts[i, j])} for j in range(tn[i])]
                for i in range(m)]

    @property
    def results(self):
        cols = [self.k, self.t2] + ([self.sigma] if self.fit != "gaussian" else [])
        if _is_torch(self.k):
            import torch
            return torch.stack(cols, dim=1).double()
        return np.stack(cols, axis=1).astype(np.float64)

    @property
    def convergence_flags(self):
        return self.status == 0

    @property
    def num_iterations(self):
        return self.nit

    @property
    def final_errors(self):
        return self.fun.double() if _is_torch(self.fun) else self.fun.astype(np.float64)

    def as_all_results(self):
        """The list ``pool.map`` returned: one ``(params, success, nit, fun, iteration_info)`` per voxel."""
        r = self.results.cpu().numpy() if _is_torch(self.k) else self.results
        ok = np.asarray(self.convergence_flags.cpu() if _is_torch(self.status) else self.convergence_flags)
        nit = np.asarray(self.nit.cpu() if _is_torch(self.nit) else self.nit)
        fun = np.asarray(self.final_errors.cpu() if _is_torch(self.fun) else self.final_errors)
        infos = self.iteration_infos
        return [(r[i], bool(ok[i]), int(nit[i]), float(fun[i]), infos[i]) for i in range(r.shape[0])]


def _device_counts(torch, device):
    """[4] zeroed int64 device counters for ONE device-memory call (t2fit_outputs.counts_dev): slot 0 = mask indices out
    of range, 1..3 = voxels per non-OK status.  Every call of the mirror owns its counters, so nothing leaks between calls
    or streams."""
    return torch.zeros(4, dtype=torch.int64, device=device)


def _raise_for_counts(cnt):
    """What the reference does for such voxels: numpy's fancy indexing raises IndexError for an index outside the array,
    scipy raises ValueError for lb > ub inside the first such voxel and pool.map aborts the whole map."""
    if cnt[0] > 0:
        raise IndexError("mask_indices out of range")
    if cnt[3] > 0:
        raise ValueError(BOUNDS_ERROR)


def resolve_solver(fit, solver):
    """``auto``: the float32 Newton/LM kernels for 'gaussian' (they reach the bounded minimiser the reference's
    ftol=1e-6 run approaches); the reference's own optimiser (L-BFGS-B restated in FP64) for 'gaussian_rician' and
    'rician', whose presets stop at ftol=gtol=1e-2, far from any minimiser -- only the same trajectory gives the
    same maps there.  Of its two forms 'auto' takes 'lbfgsb_dense' (the limited-memory matrix as the n x n matrix it
    represents: same parity with the reference on every golden fixture, 5-15x the throughput); 'lbfgsb' is scipy's
    compact 2m x 2m form restated operation by operation."""
    if solver == "auto":
        return "fast" if fit == "gaussian" else "lbfgsb_dense"
    if solver not in _abi.SOLVERS:
        raise ValueError(f"unknown solver {solver!r}")
    if solver == "fast" and fit == "rician":
        raise ValueError("fit='rician' is a negative log-likelihood, not least squares: use solver='lbfgsb_dense' or 'lbfgsb'")
    return solver


def _fill_problem(p, fit, fit_params, TEeffs, prior, norm, max_iter, tol, init_mode, solver="fast"):
    if fit not in _abi.MODELS:
        raise ValueError(f"unknown fit {fit!r}")
    p.solver = _abi.SOLVERS[solver]
    if solver in ("lbfgsb", "lbfgsb_dense"):   # fit_params['options'] as handed to scipy.optimize.minimize (:260-286)
        opt = dict(fit_params.get("options") or {})
        p.lbfgsb_ftol = float(opt.get("ftol", 0.0))
        p.lbfgsb_gtol = float(opt.get("gtol", 0.0))
        p.lbfgsb_maxls = int(opt.get("maxls", 0))
        p.lbfgsb_maxiter = int(opt.get("maxiter", 0))
        p.lbfgsb_maxfun = int(opt.get("maxfun", 0))
    x0 = list(fit_params["initial_guess"])
    bounds = list(fit_params["param_bounds"])
    npar = 2 if fit == "gaussian" else 3
    if len(x0) != npar:
        raise ValueError(f"fit {fit!r} takes {npar} parameters, initial_guess has {len(x0)}")
    if len(bounds) != len(x0):
        raise ValueError("The number of bounds is not compatible with the length of `x0`.")   # scipy's text
    te = np.ascontiguousarray(np.asarray(TEeffs, dtype=np.float64).reshape(-1))
    p.n_echo = te.size
    p.te_ms = te.ctypes.data_as(C.POINTER(C.c_double))
    p.model = _abi.MODELS[fit]
    for i in range(npar):
        lo, hi = bounds[i]
        p.x0[i] = float(x0[i])
        p.lb[i] = -np.inf if lo is None else float(lo)
        p.ub[i] = np.inf if hi is None else float(hi)
    p.no_prior = 0 if prior else 1
    p.no_prior_k_ub = NO_PRIOR_K_UB
    p.no_prior_t2_lb, p.no_prior_t2_ub = NO_PRIOR_T2_BOUNDS
    p.norm = int(bool(norm))
    p.max_iter = int(max_iter)
    p.tol = float(tol)
    if init_mode == "auto":                # 3-parameter fast solver: multi-start (several local minima); 2-parameter: log-linear
        init_mode = "best" if fit != "gaussian" else "loglinear"
    p.init = {"loglinear": _abi.INIT_LOGLINEAR, "preset": _abi.INIT_PRESET, "best": _abi.INIT_BEST}[init_mode]
    return te      # keep alive


def _run(lib, p, o, stream):
    rc = lib.t2fit_run(C.byref(p), C.byref(o), stream)
    if rc != 0:
        msg = (lib.t2fit_last_error() or b"").decode()
        if rc == -1 and "upper bound" in msg:
            raise ValueError(msg)
        if rc == -1 and "mask_idx out of range" in msg:
            raise IndexError("mask_indices out of range")
        raise _abi.T2FitError(f"t2fit_run: {_abi.ERRORS.get(rc, rc)}: {msg}")


def fit_voxels_batch(reshaped_t2w, mask_indices, TEeffs, fit, fit_params, prior=True, norm=False, *,
                     solver="auto", trace_cap=0, max_iter=0, tol=0.0, init_mode="auto", check_bounds=True,
                     dense_out=None, want=("nit", "fun", "status")) -> FitResult:
    """Fit every voxel ``mask_indices[i]`` of ``reshaped_t2w`` (float32 ``[N, E]``, run_t2mapping.py:411).

    Arguments as ``fit_voxel`` (run_t2mapping.py:120): ``fit`` in {'gaussian','gaussian_rician','rician'},
    ``fit_params`` the preset dict (``initial_guess``, ``param_bounds``, ``options``), ``prior`` False =
    ``--no_prior`` per-voxel bounds (:243-245), ``norm`` row-max normalisation (:237-240).
    ``mask_indices`` None fits all rows.  Raises ``ValueError`` where scipy would (bounds with
    lb > ub, including any masked voxel with S(TE0) > 10000 under ``--no_prior``).
    ``solver``: 'fast' | 'lbfgsb_dense' | 'lbfgsb' | 'auto' (see :func:`resolve_solver`).  ``trace_cap`` > 0 (L-BFGS-B solver)
    also returns the callback trace of every fitted voxel (``FitResult.iteration_infos``) -- meant for the
    sampled voxels of the convergence plots (utils/t2map_utils.py:115-292), pass their indices only.
    """
    lib = init()
    solver = resolve_solver(fit, solver)
    p, o = _abi.Problem(), _abi.Outputs()
    keep = [_fill_problem(p, fit, fit_params, TEeffs, prior, norm, max_iter, tol, init_mode, solver)]
    tracing = solver in ("lbfgsb", "lbfgsb_dense") and trace_cap > 0
    dev = _is_torch(reshaped_t2w)
    if dev:
        import torch
        y = reshaped_t2w
        if not y.is_cuda or y.dtype != torch.float32 or not y.is_contiguous() or y.dim() != 2:
            raise ValueError("device input must be a contiguous float32 CUDA tensor [N, E]")
        if y.device.index != _state["device"]:
            raise ValueError(f"tensor on cuda:{y.device.index}, library bound to cuda:{_state['device']}")
        n_vox, n_echo = y.shape
        idx = mask_indices
        if idx is not None:
            if not _is_torch(idx):
                idx = np.asarray(idx)
                idx = torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int32 if idx.dtype == np.int32 else np.int64), device=y.device)
            if idx.dtype not in (torch.int64, torch.int32) or not idx.is_contiguous():
                idx = idx.to(torch.int64).contiguous()
            p.idx_dtype = _abi.IDX_I32 if idx.dtype == torch.int32 else _abi.IDX_I64
        m = n_vox if idx is None else idx.numel()

        def alloc(dt=torch.float32):
            return torch.empty(m, dtype=dt, device=y.device)
        out = {"t2": alloc(), "k": alloc(), "res": alloc(),
               "sigma": alloc() if fit != "gaussian" else torch.zeros(m, dtype=torch.float32, device=y.device),
               "fun": alloc() if "fun" in want else None,
               "nit": alloc(torch.int32) if "nit" in want else None,
               "status": alloc(torch.uint8) if "status" in want else None}
        if tracing:
            out["trace_f"] = torch.full((m, trace_cap), float("nan"), dtype=torch.float32, device=y.device)
            out["trace_step"] = torch.full((m, trace_cap), float("nan"), dtype=torch.float32, device=y.device)
            out["trace_len"] = torch.zeros(m, dtype=torch.int32, device=y.device)
        p.echoes, p.memory = y.data_ptr(), _abi.MEM_DEVICE
        p.mask_idx = idx.data_ptr() if idx is not None else None
        ptr = lambda t: t.data_ptr() if t is not None else None
        stream = torch.cuda.current_stream(y.device).cuda_stream
        dcnt = _device_counts(torch, y.device)
        o.counts_dev = dcnt.data_ptr()
        keep += [y, idx, dcnt]
    else:
        y = _host_echoes(np.asarray(reshaped_t2w), p)
        if y.ndim != 2:
            raise ValueError("reshaped_t2w must be [N, E]")
        n_vox, n_echo = y.shape
        idx = None
        if mask_indices is not None:               # int32 indices are taken as they are (half the index bytes over PCIe)
            idx = np.asarray(mask_indices)
            idx = np.ascontiguousarray(idx, dtype=np.int32 if idx.dtype == np.int32 else np.int64)
            p.idx_dtype = _abi.IDX_I32 if idx.dtype == np.int32 else _abi.IDX_I64
        m = n_vox if idx is None else idx.size
        fields = [("t2", np.float32), ("k", np.float32), ("res", np.float32)]
        fields += [("sigma", np.float32)] if fit != "gaussian" else []                  # 2-parameter fit: zeros, made on first access
        fields += [(n, dt) for n, dt in (("fun", np.float32), ("nit", np.int32), ("status", np.uint8)) if n in want]
        out = {"sigma": None, "fun": None, "nit": None, "status": None}
        out.update(_host_block(m, fields))
        if tracing:
            out["trace_f"] = np.full((m, trace_cap), np.nan, np.float32)
            out["trace_step"] = np.full((m, trace_cap), np.nan, np.float32)
            out["trace_len"] = np.zeros(m, np.int32)
        p.echoes, p.memory = y.ctypes.data, _abi.MEM_HOST
        p.mask_idx = idx.ctypes.data if idx is not None else None
        ptr = lambda a: a.ctypes.data if a is not None else None
        stream = None
        keep += [y, idx]
    if n_echo != p.n_echo:
        raise ValueError(f"reshaped_t2w has {n_echo} echoes, TEeffs has {p.n_echo}")
    if tracing:
        o.trace_f, o.trace_step, o.trace_len = ptr(out["trace_f"]), ptr(out["trace_step"]), ptr(out["trace_len"])
        o.trace_cap = int(trace_cap)
    p.layout, p.ld, p.n_vox, p.n_fit = _abi.LAYOUT_AOS, 0, n_vox, m
    o.t2, o.k, o.res = ptr(out["t2"]), ptr(out["k"]), ptr(out["res"])
    o.sigma = ptr(out["sigma"]) if fit != "gaussian" else None
    o.fun, o.nit, o.status = ptr(out["fun"]), ptr(out["nit"]), ptr(out["status"])
    o.dense = 0
    _run(lib, p, o, stream)
    if dev:
        counts = None
        if check_bounds:                           # one small D2H: waits for the call (check_bounds=False stays asynchronous)
            cnt = [int(v) for v in dcnt.cpu()]
            _raise_for_counts(cnt)
            counts = (m - cnt[1] - cnt[2] - cnt[3], cnt[1], cnt[2], cnt[3])
    else:
        counts = tuple(o.status_count)
        if counts[3] > 0:
            # scipy raises inside the first such voxel and the reference's pool.map aborts the whole map
            raise ValueError(BOUNDS_ERROR)
    return FitResult(out["t2"], out["k"], out["sigma"], out["res"], out["fun"], out["nit"], out["status"], fit,
                     counts if counts is not None else (0, 0, 0, 0), solver, out.get("trace_f"), out.get("trace_step"),
                     out.get("trace_len"))


def fit_voxels_into(reshaped_t2w, mask_indices, TEeffs, fit, fit_params, prior, norm, out, *, solver="auto", counts=None):
    """Device-memory fit whose compact results go to caller-given raw device pointers: ``out`` maps 't2', 'k', 'res'
    (and 'sigma' for the 3-parameter fits; optionally 'status', 'nit', 'fun') to integer addresses of arrays with room
    for ``len(mask_indices)`` elements -- local memory or a peer GPU's buffer mapped with ``t2fit_shared_open`` (the
    fused gather of ``distributed.fit_voxels_sharded``).  Asynchronous on the current torch stream.  ``counts``: a zeroed
    int64 CUDA tensor [4] the call adds its histogram to (slot 0 indices out of range, 1..3 voxels per non-OK status);
    the caller reads it when it synchronises and raises as :func:`fit_voxels_batch` does (``check_counts``)."""
    import torch
    lib = init()
    solver = resolve_solver(fit, solver)
    p, o = _abi.Problem(), _abi.Outputs()
    keep = [_fill_problem(p, fit, fit_params, TEeffs, prior, norm, 0, 0.0, "auto", solver)]
    y = reshaped_t2w
    if not (_is_torch(y) and y.is_cuda and y.dtype == torch.float32 and y.is_contiguous() and y.dim() == 2):
        raise ValueError("device input must be a contiguous float32 CUDA tensor [N, E]")
    idx = mask_indices
    if idx is not None and not (_is_torch(idx) and idx.is_cuda and idx.dtype in (torch.int64, torch.int32) and idx.is_contiguous()):
        idx = torch.as_tensor(np.ascontiguousarray(np.asarray(idx.cpu() if _is_torch(idx) else idx), dtype=np.int64), device=y.device)
    if idx is not None:
        p.idx_dtype = _abi.IDX_I32 if idx.dtype == torch.int32 else _abi.IDX_I64
    if counts is None:
        counts = _device_counts(torch, y.device)           # never the per-process counters: nothing leaks into later calls
    o.counts_dev = counts.data_ptr()
    if y.shape[1] != p.n_echo:
        raise ValueError(f"reshaped_t2w has {y.shape[1]} echoes, TEeffs has {p.n_echo}")
    p.echoes, p.memory, p.layout = y.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS
    p.mask_idx = idx.data_ptr() if idx is not None else None
    p.n_vox, p.n_fit = y.shape[0], (y.shape[0] if idx is None else idx.numel())
    o.t2, o.k, o.res = out["t2"], out["k"], out["res"]
    o.sigma = out.get("sigma") if fit != "gaussian" else None
    o.status, o.nit, o.fun = out.get("status"), out.get("nit"), out.get("fun")
    o.dense = 0
    _run(lib, p, o, torch.cuda.current_stream(y.device).cuda_stream)
    keep += [y, idx, counts]
    return p.n_fit


def check_counts(counts):
    """Raise what the reference raises for the histogram of a device call (synchronises on the tensor)."""
    _raise_for_counts([int(v) for v in counts.cpu()])


def mask_indices_device(mask, n_masks=None):
    """``mask_indices`` of a device mask: union over the per-TE axis (``np.sum(mask, axis=3) > 0``,
    run_t2mapping.py:383-384) and ascending C-order compaction (:412,:421).  ``mask`` is a uint8/bool
    CUDA tensor ``[z,y,x]`` or ``[z,y,x,n_masks]``; returns an int64 CUDA tensor."""
    import torch
    lib = init()
    mk = mask
    if mk.dtype == torch.bool:
        mk = mk.view(torch.uint8)
    if mk.dtype != torch.uint8:
        mk = (mk > 0).view(torch.uint8)
    mk = mk.contiguous()
    nm = 1 if mk.dim() == 3 else mk.shape[-1]
    n_vox = mk.numel() // nm
    idx = torch.empty(n_vox, dtype=torch.int64, device=mk.device)
    n = C.c_int64()
    stream = torch.cuda.current_stream(mk.device).cuda_stream
    _abi.check(lib, lib.t2fit_mask_indices(mk.data_ptr(), n_vox, nm, idx.data_ptr(), C.byref(n), stream),
               "t2fit_mask_indices")
    return idx[:n.value]


def t2map_volume(t2w, mask, TEeffs, fit, fit_params, prior=True, norm=False, **kw):
    """Stacked per-TE volumes ``t2w[z,y,x,E]``, mask ``[z,y,x]`` or per-TE masks ``[z,y,x,E]`` and the
    TE vector in; ``(t2_map, k_map, sigma_map, res_map)`` float32 ``[z,y,x]`` out, zeros off-mask --
    the whole hot block of ``process_t2maps`` (run_t2mapping.py:383-386, :411-461, :471-473)."""
    lib = init()
    shape3 = tuple(t2w.shape[:3])
    n_echo = t2w.shape[-1]
    p, o = _abi.Problem(), _abi.Outputs()
    solver = resolve_solver(fit, kw.get("solver", "auto"))
    keep = [_fill_problem(p, fit, fit_params, TEeffs, prior, norm, kw.get("max_iter", 0), kw.get("tol", 0.0),
                          kw.get("init_mode", "auto"), solver)]
    fused_mask = None
    if _is_torch(t2w):
        import torch
        y = t2w.reshape(-1, n_echo)
        if y.dtype != torch.float32:
            y = y.float()
        y = y.contiguous()
        mk = mask if mask.dim() == 3 else (mask != 0).any(dim=3)                       # :383-384
        mk = (mk if mk.dtype in (torch.bool, torch.uint8) else (mk != 0)).contiguous()
        mk = mk.view(torch.uint8) if mk.dtype == torch.bool else mk
        idx = mask_indices_device(mk)                                                   # :412,:421
        n_vox, m = y.shape[0], idx.numel()
        # the four maps are zero-filled by the fit launch itself (or zero_fill_kernel on a side stream), see t2fit.h (:415-418)
        maps = torch.empty((4, n_vox), dtype=torch.float32, device=y.device)
        fused_mask = mk.reshape(-1)
        p.echoes, p.memory, p.mask_idx = y.data_ptr(), _abi.MEM_DEVICE, idx.data_ptr()
        stream = torch.cuda.current_stream(y.device).cuda_stream
        mp = [maps[i].data_ptr() for i in range(4)]
        dcnt = _device_counts(torch, y.device)
        o.counts_dev = dcnt.data_ptr()
        keep += [y, idx, mk, dcnt]
    else:
        y = _host_echoes(np.reshape(np.asarray(t2w), (-1, n_echo)), p)                  # :411
        mk = np.asarray(mask)
        mk = (np.sum(mk, axis=3) > 0) if mk.ndim == 4 else (mk if mk.dtype == np.bool_ else mk > 0)    # :383-384
        idx = np.flatnonzero(mk.reshape(-1)).astype(np.int64, copy=False)               # :412,:421
        n_vox, m = y.shape[0], idx.size
        maps = np.zeros((4, n_vox), np.float32)                                         # :415-418
        p.echoes, p.memory, p.mask_idx = y.ctypes.data, _abi.MEM_HOST, idx.ctypes.data
        stream = None
        mp = [maps[i].ctypes.data for i in range(4)]
        keep += [y, idx]
    if n_echo != p.n_echo:
        raise ValueError(f"t2w has {n_echo} echoes, TEeffs has {p.n_echo}")
    p.layout, p.ld, p.n_vox, p.n_fit = _abi.LAYOUT_AOS, 0, n_vox, m
    o.t2, o.k, o.res = mp[0], mp[1], mp[3]
    o.sigma = mp[2] if (fit != "gaussian" or fused_mask is not None) else None
    o.dense = 1
    if fused_mask is not None:
        o.zero_fill_mask = fused_mask.data_ptr()
    _run(lib, p, o, stream)
    if stream is not None:
        _raise_for_counts([int(v) for v in dcnt.cpu()])
    elif o.status_count[3] > 0:
        raise ValueError(BOUNDS_ERROR)
    return tuple(maps[i].reshape(shape3) for i in range(4))


def compute_residuals(reshaped_t2w, TEeffs, fit, norm, k_map, t2_map, sigma_map, res_map, mask_indices, mask):
    """Signature of the reference's ``compute_residuals`` (utils/t2map_utils.py:62-89) for callers that
    keep the two-step structure (``fit_voxels_batch`` already returns the same residuals).  Runs the
    stand-alone CUDA residual pass (``t2fit_residuals``); numpy inputs are staged through torch."""
    import torch
    lib = init()
    dev = torch.device("cuda", _state["device"])
    host = not _is_torch(reshaped_t2w)

    def to_dev(a, dt):
        t = a if _is_torch(a) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(device=dev, dtype=dt).contiguous()
    y = to_dev(reshaped_t2w, torch.float32)
    idx = to_dev(mask_indices, torch.int64)
    k, t2, sg = (to_dev(a, torch.float32).reshape(-1) for a in (k_map, t2_map, sigma_map))
    res = to_dev(res_map, torch.float32).reshape(-1).clone()
    p = _abi.Problem()
    dummy = {"initial_guess": [1.0, 100.0] + ([1.0] if fit != "gaussian" else []),
             "param_bounds": [(0.0, 1e9), (1e-3, 1e9)] + ([(0.0, 1e9)] if fit != "gaussian" else [])}
    keep = _fill_problem(p, fit, dummy, TEeffs, True, norm, 0, 0.0, "loglinear")
    p.echoes, p.memory, p.layout, p.mask_idx = y.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS, idx.data_ptr()
    p.n_vox, p.n_fit = y.shape[0], idx.numel()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _abi.check(lib, lib.t2fit_residuals(C.byref(p), k.data_ptr(), t2.data_ptr(), sg.data_ptr(), res.data_ptr(), stream),
               "t2fit_residuals")
    del keep
    shape3 = tuple(mask.shape[:3])
    if host:
        return res.cpu().numpy().reshape(shape3)
    return res.reshape(shape3)
