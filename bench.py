#!/usr/bin/env python
"""Headline benchmark: masked-voxel T2 fits per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- adult-brain 256x256x256 x 5 TE, ellipsoid
brain mask (~1.6 M masked voxels), 2-parameter mono-exponential fit, LF preset, --no_prior.
One *step* = the whole hot block of process_t2maps for one such volume (run_t2mapping.py:411-461):
zero the four dense maps, fit every masked voxel, residual epilogue, scatter into the dense maps,
plus convergence flags / iteration counts / final errors per voxel.

  value     whole-job fits/s with the volume already resident in HBM (AoS [N,E] float32 + mask_indices)
  e2e       same metric through fit_voxels_batch() with HOST (page-locked) numpy buffers, host<->device traffic inside
  roofline  the step launch (fit + zero-fill of the dense maps in one kernel) against the measured HBM peak: algorithmic
            bytes / step time from CUDA events over the timed region; roofline_fp32: the plain fit launch alone
            (events around each launch) against the FP32 / MUFU peaks
  solver_lbfgsb / solver_floor3_fast / parity   secondary blocks outside the timed region
  cpu_baseline  the oracle port (scipy L-BFGS-B exactly as the reference drives it) on a bounded sample
N > 1: weak scaling, one volume-sized slab of masked voxels per rank, no data-path collective; the
final NCCL gather of the parameter maps is timed separately ("final_gather").

--impl reference: the reference's own CPU implementation of the path (oracle port: scipy.optimize
L-BFGS-B + multiprocessing over all host cores), each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os

# the CPU arm forks scipy workers: BLAS must be single-threaded BEFORE numpy loads it (SURVEY.md section 6:
# without this the forked workers oversubscribe the cores and the baseline collapses 50x)
for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "c2: 256x256x256 x 5 TE adult-brain, ellipsoid mask, gaussian 2-param fit, LF preset, --no_prior"
METRIC = "masked_voxel_T2_fits_per_sec"
UNIT = "fits/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)), "src": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def make_workload(rank):
    from fetal_t2mapping_b200 import synth
    y, mask, te, _ = synth.make_volume("c2", scale=float(os.environ.get("T2FIT_BENCH_SCALE", "1.0")), volume_index=rank)
    flat = np.ascontiguousarray(y.reshape(-1, te.size))
    idx = np.flatnonzero(mask.reshape(-1)).astype(np.int64)
    return flat, idx, te


def cpu_sample(flat, idx, te, fp, n, seed=11):
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(idx, size=min(n, idx.size), replace=False))
    return np.ascontiguousarray(flat[pick])


_last_nit = [None]


def time_oracle(rows, te, fp, procs):
    from oracle import fit_oracle as fo
    t0 = time.perf_counter()
    p, ok, nit, fun, _ = fo.fit_rows_oracle(rows, te, "gaussian", fp, False, False, mode="verbatim", procs=procs)
    dt = time.perf_counter() - t0
    _last_nit[0] = nit
    return rows.shape[0] / dt, dt, p, ok


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: scipy L-BFGS-B driven exactly as fit_voxel drives it, Pool over all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import fit_oracle as fo
    _, fp = fo.preset("gaussian", "lf")
    flat, idx, te = make_workload(0)
    procs = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    probe = cpu_sample(flat, idx, te, fp, 256 * min(procs, 8), seed=5)
    rate, _, _, _ = time_oracle(probe, te, fp, procs)
    budget_s = float(os.environ.get("T2FIT_REF_BUDGET_S", "100"))
    per_step = int(np.clip(budget_s * rate / max(1, args.steps + args.warmup), 256, 20000))
    for w in range(args.warmup):
        time_oracle(cpu_sample(flat, idx, te, fp, per_step, seed=100 + w), te, fp, procs)
    t_tot, n_tot = 0.0, 0
    for s in range(args.steps):
        rows = cpu_sample(flat, idx, te, fp, per_step, seed=200 + s)
        _, dt, _, _ = time_oracle(rows, te, fp, procs)
        t_tot += dt
        n_tot += rows.shape[0]
    value = n_tot / t_tot
    import scipy
    sample = f"{per_step} masked voxels per step (seeded random sample of the {idx.size}-voxel mask), {args.steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "masked_voxels": int(idx.size), "sample": sample,
                       "scipy": scipy.__version__, "numpy": np.__version__},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # CPU baseline first: multiprocessing fork must happen before CUDA is initialised in this process
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import fit_oracle as fo
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        _, fp0 = fo.preset("gaussian", "lf")
        flat0, idx0, te0 = make_workload(0)
        procs = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        probe_rate, _, _, _ = time_oracle(cpu_sample(flat0, idx0, te0, fp0, 64 * procs, seed=5), te0, fp0, procs)
        n_s = int(os.environ.get("T2FIT_CPU_SAMPLE", "0")) or int(np.clip(20.0 * probe_rate, 512, 20000))
        rows = cpu_sample(flat0, idx0, te0, fp0, n_s)
        rate, dt, cpu_params, cpu_ok = time_oracle(rows, te0, fp0, procs)
        cpu_rows, cpu_nit = rows.copy(), _last_nit[0].copy()
        import scipy
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                        "sample": f"{rows.shape[0]} seeded random masked voxels of the same volume, {dt:.1f} s, "
                                  f"scipy {scipy.__version__} L-BFGS-B via multiprocessing.Pool({procs})"}
        del flat0, idx0, rows
    else:
        cpu_rows = cpu_params = cpu_ok = cpu_nit = None
    import torch
    import torch.distributed as dist
    import fetal_t2mapping_b200 as t2
    from fetal_t2mapping_b200 import _abi
    import ctypes as C

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = t2.init(local)
    _, fp = t2.preset("gaussian", True)

    flat, idx, te = make_workload(rank)
    n_vox, n_echo = flat.shape
    m = idx.size
    y_d = torch.from_numpy(flat).to(dev)
    idx_d = torch.from_numpy(idx).to(dev)
    maps = torch.empty((4, n_vox), dtype=torch.float32, device=dev)
    mask_np = np.zeros(n_vox, np.uint8)
    mask_np[idx] = 1
    mask_d = torch.from_numpy(mask_np).to(dev)          # the (union) mask volume, reshaped_mask of :412
    fun_d = torch.empty(m, dtype=torch.float32, device=dev)
    nit_d = torch.empty(m, dtype=torch.int32, device=dev)
    st_d = torch.empty(m, dtype=torch.uint8, device=dev)

    p, o = _abi.Problem(), _abi.Outputs()
    from fetal_t2mapping_b200.api import _fill_problem
    keep = _fill_problem(p, "gaussian", fp, te, False, False, 0, 0.0, "loglinear")
    p.echoes, p.memory, p.layout, p.mask_idx = y_d.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS, idx_d.data_ptr()
    p.n_vox, p.n_fit = n_vox, m
    o.t2, o.k, o.sigma, o.res = maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), maps[3].data_ptr()
    o.fun, o.nit, o.status, o.dense = fun_d.data_ptr(), nit_d.data_ptr(), st_d.data_ptr(), 1
    fused = os.environ.get("T2FIT_BENCH_FUSED_FILL", "1") == "1"
    # zero-fill route of the library: inside the fit launch (default) or zero_fill_kernel on a side stream (T2FIT_FILL=stream)
    fill_in_fit = fused and os.environ.get("T2FIT_FILL", "fused") != "stream"
    if fused:
        o.zero_fill_mask = mask_d.data_ptr()             # np.zeros_like x4 (:415-418) done by the fit launch itself
    stream = torch.cuda.current_stream(dev)

    def step(ev=None):
        if not fused:
            maps.zero_()                               # np.zeros_like x4 (run_t2mapping.py:415-418)
        if ev is not None:
            ev[0].record(stream)
        rc = lib.t2fit_run(C.byref(p), C.byref(o), stream.cuda_stream)
        if ev is not None:
            ev[1].record(stream)
        if rc:
            raise RuntimeError(lib.t2fit_last_error())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    # the timed region is a few ms: keep the same load running ~1.2 s before it (untimed) so that clocks
    # have settled and nvidia-smi (100 ms period) has samples under load
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < float(os.environ.get("T2FIT_BENCH_SPIN_S", "1.2")):
        for _ in range(50):
            step()
        torch.cuda.synchronize()
    barrier()
    # the timed region: K steps between two CUDA events on the launching stream.  No per-step events inside:
    # an event pair around every fork/join step was measured to stretch a 68 us step to 125 us.
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for s in range(args.steps):
        step()
    t_end.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = t_start.elapsed_time(t_end)
    kern_ms = elapsed_ms / args.steps
    if world > 1:
        t = torch.tensor([elapsed_ms, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, kern_ms = float(t[0]), float(t[1])
        tot = torch.tensor([m], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        m_total = int(tot[0])
    else:
        m_total = m
    value = m_total * args.steps / (elapsed_ms * 1e-3)

    # the fit kernel alone (the zero-fill runs concurrently inside a step and would blur an event pair
    # around it): same launch, zero_fill_mask off, CUDA events around each launch, K launches
    o_fit = _abi.Outputs()
    o_fit.t2, o_fit.k, o_fit.sigma, o_fit.res = o.t2, o.k, None, o.res
    o_fit.fun, o_fit.nit, o_fit.status, o_fit.dense = o.fun, o.nit, o.status, 1
    for _ in range(3):
        lib.t2fit_run(C.byref(p), C.byref(o_fit), stream.cuda_stream)
    torch.cuda.synchronize()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a_, b_ in kev:
        a_.record(stream)
        lib.t2fit_run(C.byref(p), C.byref(o_fit), stream.cuda_stream)
        b_.record(stream)
    torch.cuda.synchronize()
    fit_ms = float(np.mean([a_.elapsed_time(b_) for a_, b_ in kev]))
    if world > 1:
        t = torch.tensor([fit_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fit_ms = float(t[0])

    # sanity of the timed result + work actually done (passes per voxel) for the roofline numerators
    nit = nit_d.cpu().numpy().astype(np.int64)
    status = st_d.cpu().numpy()
    n_failed = int((status != 0).sum())
    assert n_failed <= 1e-5 * m, f"bench workload produced {n_failed} failed voxels"
    t2v = maps[0][idx_d].cpu().numpy()
    assert np.isfinite(t2v).all() and t2v.min() >= 10 and t2v.max() <= 2000
    off_mask_ok = bool((maps[:, mask_d == 0] == 0).all()) if fused else None
    assert off_mask_ok in (True, None), "dense maps are not zero off-mask"
    wm = t2.work_model("gaussian", n_echo)
    passes = nit                                        # passes over the echoes the solver needed, per voxel
    flops_launch = float((wm["flop_fixed"] + wm["flop_per_pass"] * passes).sum())
    mufu_launch = float((wm["mufu_fixed"] + wm["mufu_per_pass"] * passes).sum())
    fit_bytes = float(m) * (wm["bytes_per_voxel"] + 8 + 4 + 4)          # + int64 index, nit, fun
    # whole step: + mask bytes read, zeros written to every unmasked slot of t2/k/res and to all of sigma
    step_bytes = fit_bytes + (float(n_vox) + 4.0 * (3 * (n_vox - m) + n_vox) if fused else 4.0 * 4 * n_vox)
    peaks = measured_peaks()
    info = t2.device_info()
    fp32_peak = info["sm_count"] * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12      # TFLOP/s at max clock
    mufu_peak = info["sm_count"] * 16 * peaks["sm_max_mhz"] * 1e6 / 1e12
    ach_tf = flops_launch / (fit_ms * 1e-3) / 1e12
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        traffic = json.load(open(tp))
    roof_fp32 = {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                 "traffic": traffic.get("fit_kernel_dram_bytes_per_launch"),
                 "peak_src": f"148 SM x 128 lanes x 2 x sm_max_mhz ({peaks['src']} clocks)",
                 "kernel": "fit_kernel<mono2,E=5,AoS> alone", "kernel_ms": fit_ms,
                 "mufu_frac": mufu_launch / (fit_ms * 1e-3) / 1e12 / mufu_peak,
                 "hbm_gbs": fit_bytes / (fit_ms * 1e-3) / 1e9,
                 "passes_per_voxel": float(passes.mean()), "flop_per_voxel": flops_launch / m}
    ach_gbs = step_bytes / (kern_ms * 1e-3) / 1e9
    roof_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach_gbs / peaks["hbm_gbs"], "traffic": traffic.get("step_dram_bytes"),
                "peak_src": peaks["src"],
                "kernel": ("fit_kernel<mono2,E=5,AoS,FILL> (one launch per step: fit + zero-fill of the dense maps)" if fill_in_fit
                           else "fit_kernel || zero_fill_kernel (one step, concurrent streams)"),
                "kernel_ms": kern_ms, "algorithmic_bytes": step_bytes}
    # the step is bound by HBM (the four dense float32 maps are 268 MB of mostly zeros); the fit kernel by FP32/MUFU
    roofline = dict(roof_hbm)

    # end to end through the public API with HOST buffers (numpy in, numpy out), every step: inputs from page-locked host
    # memory to the GPU, fit, results back into host arrays
    def time_e2e(a_flat, a_idx, steps, windows=3):
        # warm for >= 0.5 s: pinned result blocks cached, staging threads awake, host cores out of their idle states (a
        # 5-call warm-up left the first window 2x slower on some boxes); then `windows` timed windows of `steps` calls
        t_w, n_w = time.perf_counter(), 0
        while n_w < 5 or time.perf_counter() - t_w < 0.5:
            r_ = t2.fit_voxels_batch(a_flat, a_idx, te, "gaussian", fp, prior=False, norm=False)
            n_w += 1
        dts = []
        for _ in range(windows):
            barrier()
            t0_ = time.perf_counter()
            for _ in range(steps):
                r_ = t2.fit_voxels_batch(a_flat, a_idx, te, "gaussian", fp, prior=False, norm=False)
                _ = float(r_.res[0])
            torch.cuda.synchronize()
            dt_ = time.perf_counter() - t0_
            if world > 1:
                t_ = torch.tensor([dt_], device=dev, dtype=torch.float64)
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
                dt_ = float(t_[0])
            dts.append(dt_)
        return float(np.median(dts)), r_, dts

    e2e_steps = max(3, min(args.steps, 20))
    flat_p, idx_p = t2.pinned_array(None, like=flat), t2.pinned_array(None, like=idx)
    e2e_s, r, e2e_windows = time_e2e(flat_p, idx_p, e2e_steps)
    assert np.array_equal(r.t2, t2v), "e2e path and device path disagree"
    mapped = os.environ.get("T2FIT_HOST_IN", "auto") != "staged"      # page-locked arrays: no staging (run_host_mapped)
    e2e = {"value": m_total * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(m * (n_echo * 4 + 8)),
           "d2h_bytes_per_step": int(m * (4 * 4 + 4 + 1)), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "windows_ms_per_step": [round(1e3 * w / e2e_steps, 4) for w in e2e_windows],     # value = the median window
           "path": "fit_voxels_batch(page-locked numpy [N,E], mask_indices) -> numpy results: " +
                   ("ONE kernel gathers the masked rows (and the index vector) straight from host memory over PCIe and stores "
                    "the results straight back into the page-locked numpy result arrays; no staging, no host thread touches the data" if mapped else
                    "threaded gather into pinned staging, H2D per 2.6 MB chunk, fit kernel storing results straight into the "
                    "page-locked numpy result arrays (zero-copy D2H)")}
    # the same call with the pageable arrays a drop-in caller has (np.reshape(...).astype(np.float32), np.where)
    pg_s, r, _ = time_e2e(flat, idx, e2e_steps)
    assert np.array_equal(r.t2, t2v)
    e2e["pageable_input"] = {"value": m_total * e2e_steps / pg_s, "ms_per_step": 1e3 * pg_s / e2e_steps}
    del flat_p, idx_p

    # the reference-faithful solver (FP64 L-BFGS-B, T2FIT_SOLVER_LBFGSB) on the same device-resident workload: secondary
    # number, outside the timed region; also cross-checks the two solvers against each other
    lbfgsb = None
    if not args.no_lbfgsb:
        try:
            for _ in range(2):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                rl = t2.fit_voxels_batch(y_d, idx_d, te, "gaussian", fp, prior=False, norm=False, solver="lbfgsb", check_bounds=False)
                g1.record(stream)
                torch.cuda.synchronize()
            lb_ms = g0.elapsed_time(g1)
            t2l = rl.t2.cpu().numpy()
            lbfgsb = {"fits_per_s": m / (lb_ms * 1e-3), "ms_per_volume": lb_ms, "mean_nit": float(rl.nit.float().mean()),
                      "success": float((rl.status == 0).float().mean()),
                      "t2_within_1e-3_of_fast_solver": float(np.mean(np.abs(t2l - t2v) <= 1e-3 * np.abs(t2v))),
                      "kernel": "lbfgsb_kernel<gaussian> (one thread per voxel, FP64, state in local memory)", "dtype": "f64"}
        except Exception as ex:              # secondary number: never lose the headline line over it
            lbfgsb = {"error": repr(ex)}

    # the 3-parameter fast solver (floor_queue_kernel) on a slab of BASELINE config 5 (unmasked, 16 TE, Rician data made on the
    # device): secondary number, outside the timed region, rank 0 only
    floor3 = None
    if not args.no_lbfgsb and rank == 0:
        try:
            n5, e5 = 1 << 22, 16
            te5 = np.linspace(100, 700, e5)
            g5 = torch.Generator(device=dev).manual_seed(4)
            t5 = torch.exp(torch.empty(n5, device=dev).uniform_(np.log(10.0), np.log(2000.0), generator=g5))
            a5 = torch.empty(n5, device=dev).uniform_(300.0, 3000.0, generator=g5)
            s5 = a5[:, None] * torch.exp(-torch.tensor(te5, device=dev, dtype=torch.float32)[None, :] / t5[:, None])
            y5 = torch.sqrt((s5 + torch.randn((n5, e5), device=dev, generator=g5) * 20.0) ** 2 +
                            (torch.randn((n5, e5), device=dev, generator=g5) * 20.0) ** 2)
            del s5, a5
            fp5 = t2.preset("gaussian_rician", True)[1]
            best5 = 1e30
            for _ in range(3):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                r5 = t2.fit_voxels_batch(y5, None, te5, "gaussian_rician", fp5, prior=False, norm=False, solver="fast", check_bounds=False)
                g1.record(stream)
                torch.cuda.synchronize()
                best5 = min(best5, g0.elapsed_time(g1))
            ok5 = r5.status == 0
            rel5 = (r5.t2[ok5] - t5[ok5]).abs() / t5[ok5]
            floor3 = {"fits_per_s": n5 / (best5 * 1e-3), "ms": best5, "voxels": n5, "n_echo": e5,
                      "mean_accepted_iterations": float(r5.nit.float().mean()), "not_converged": float((~ok5).float().mean()),
                      "median_rel_error_vs_true_t2": float(rel5.median()),
                      "workload": "slab of BASELINE config 5: unmasked, 16 TE 100..700 ms, T2 log-uniform 10..2000 ms, Rician sigma 20, "
                                  "gaussian_rician LF preset --no_prior",
                      "kernel": "floor_queue_kernel<16,AoS> (persistent grid, lanes pull voxels from a queue)", "dtype": "f32"}
            del y5, t5, r5
        except Exception as ex:              # secondary number: never lose the headline line over it
            floor3 = {"error": repr(ex)}

    # Delta-T2 against the reference's scipy fit (BASELINE metric): the voxels the CPU arm just fitted, refitted by both CUDA solvers
    parity = None
    if cpu_rows is not None:
        try:
            ref_t2 = cpu_params[:, 1]
            parity = {"sample": int(cpu_rows.shape[0]), "reference": "oracle port = scipy L-BFGS-B exactly as fit_voxel calls it",
                      "reference_success": float(np.mean(cpu_ok))}
            for name in ("fast", "lbfgsb"):
                rr = t2.fit_voxels_batch(cpu_rows, None, te, "gaussian", fp, prior=False, norm=False, solver=name)
                rel = np.abs(rr.t2.astype(np.float64) - ref_t2) / np.abs(ref_t2)
                parity[name] = {"t2_rel_le_1e-3": float(np.mean(rel <= 1e-3)), "t2_rel_median": float(np.median(rel)),
                                "t2_rel_p999": float(np.quantile(rel, 0.999)), "success_equal": bool(np.array_equal(rr.status == 0, cpu_ok))}
                if name == "lbfgsb":
                    parity[name]["nit_equal"] = float(np.mean(rr.nit == cpu_nit)) if cpu_nit is not None else None
        except Exception as ex:
            parity = {"error": repr(ex)}

    # final gather of the parameter maps (north_star: the only inter-GPU traffic), timed on its own
    final_gather = None
    if world > 1:
        from fetal_t2mapping_b200 import distributed as D
        loc = torch.stack([maps[0][idx_d], maps[1][idx_d], maps[3][idx_d]]).contiguous()
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = m
        dist.all_reduce(sizes)
        cuts = [0] + torch.cumsum(sizes, 0).tolist()
        bounds = [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]
        for _ in range(3):
            full = D.gather_slabs(loc, bounds)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        full = D.gather_slabs(loc, bounds)
        g1.record(stream)
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        a0, b0 = bounds[rank]
        final_gather = {"op": "nccl all_gather_into_tensor of the (t2,k,res) slabs + stitch", "ms": float(gms[0]),
                        "bytes_per_rank": int(3 * int(sizes.max()) * 4), "gathered_voxels": int(cuts[-1]),
                        "checksum_ok": bool(torch.equal(full[:, a0:b0], loc))}

    # the same gather FUSED into the fit kernels: every rank's kernel stores its compact (t2, k, res, status) results straight
    # into rank 0's buffer over NVLink (CUDA-IPC mapping, peer stores from the epilogue), no collective
    fused_gather = None
    if world > 1:
        from fetal_t2mapping_b200.api import fit_voxels_into
        sizes_l = [int(v) for v in sizes.tolist()]
        tot_m = sum(sizes_l)
        nbytes = 13 * tot_m
        ptr, payload = C.c_void_p(), [None]
        ok_local = 1
        if rank == 0:
            hbuf = C.create_string_buffer(64)
            if lib.t2fit_shared_alloc(nbytes, C.byref(ptr), hbuf) != 0:
                ok_local = 0
            payload = [hbuf.raw if ok_local else None]
        dist.broadcast_object_list(payload, src=0)
        if rank != 0 and (payload[0] is None or lib.t2fit_shared_open(payload[0], C.byref(ptr)) != 0):
            ok_local = 0
        flag = torch.tensor([ok_local], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if world > 1 and int(flag[0]) == 0:                  # no peer mapping on this box: every rank skips consistently
        if ptr.value:
            (lib.t2fit_shared_free if rank == 0 else lib.t2fit_shared_close)(ptr)
        fused_gather = {"unavailable": (lib.t2fit_last_error() or b"").decode() or "CUDA IPC peer mapping failed on a rank"}
    elif world > 1:
        base, a0 = ptr.value, int(cuts[rank])
        outp = {"t2": base + 4 * a0, "k": base + 4 * (tot_m + a0), "res": base + 4 * (2 * tot_m + a0), "status": base + 12 * tot_m + a0}
        for _ in range(3):
            fit_voxels_into(y_d, idx_d, te, "gaussian", fp, False, False, outp, solver="fast")
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            fit_voxels_into(y_d, idx_d, te, "gaussian", fp, False, False, outp, solver="fast")
        f1.record(stream)
        barrier()
        fms = torch.tensor([f0.elapsed_time(f1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(fms, op=dist.ReduceOp.MAX)
        ok = None
        if rank == 0:
            class _Raw:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (base, False), "version": 2}
            raw = torch.as_tensor(_Raw(), device=dev)
            got = raw[:4 * tot_m].view(torch.float32)
            ok = bool(torch.equal(got[:m], maps[0][idx_d])) and bool(torch.equal(got[cuts[1]:cuts[1] + 8], full[0, cuts[1]:cuts[1] + 8]))
            del raw, got
        barrier()
        (lib.t2fit_shared_free if rank == 0 else lib.t2fit_shared_close)(ptr)
        fused_gather = {"op": "fit_kernel epilogue stores (t2,k,res,status) into rank 0's buffer over NVLink (CUDA IPC peer mapping)",
                        "ms_per_step_fit_plus_gather": float(fms[0]), "local_fit_ms": fit_ms, "nccl_gather_ms": final_gather["ms"],
                        "bytes_to_root_per_rank": int(13 * m), "checksum_ok": ok}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "masked_voxels_per_gpu": int(m), "failed_voxels": n_failed, "volume_voxels_per_gpu": int(n_vox),
                           "n_echo": int(n_echo), "l2": "inputs+outputs per step (603 MB) exceed the 126 MB L2",
                           "partition": f"weak: one volume-sized slab per rank x {world}", "gather": "none in the timed step",
                           "zero_fill": ("inside the fit launch: every fit thread zeroes a few 4-voxel words of the dense maps" if fill_in_fit else
                                         "zero_fill_kernel on a side stream, concurrent with fit_kernel" if fused else "torch zero_() before the fit launch")},
                "roofline": roofline, "roofline_fp32": roof_fp32, "roofline_hbm": roof_hbm, "cpu_baseline": cpu_baseline,
                "solver_lbfgsb": lbfgsb, "solver_floor3_fast": floor3, "parity": parity, "e2e": e2e, "gpu_launches": int(args.steps) * (2 if (fused and not fill_in_fit) else 1), "clocks": clocks, "final_gather": final_gather, "fused_gather": fused_gather,
                "device": info["name"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    del keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-lbfgsb", dest="no_lbfgsb", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
