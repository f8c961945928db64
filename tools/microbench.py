"""Component timings on the GPU box (CUDA events, warm, 50 reps each): where does a step's time go?"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fetal_t2mapping_b200 as t2
from fetal_t2mapping_b200 import _abi
from fetal_t2mapping_b200.api import _fill_problem
from bench import make_volume_workload


def make_workload(rank):
    return make_volume_workload("c2", rank)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3     # us


def main():
    lib = t2.init(0)
    flat, idx, te = make_workload(0)
    n_vox, n_echo = flat.shape
    m = idx.size
    dev = torch.device("cuda", 0)
    y_d, idx_d = torch.from_numpy(flat).to(dev), torch.from_numpy(idx).to(dev)
    mask_np = np.zeros(n_vox, np.uint8); mask_np[idx] = 1
    mask_d = torch.from_numpy(mask_np).to(dev)
    maps = torch.empty((4, n_vox), dtype=torch.float32, device=dev)
    comp = torch.empty((4, m), dtype=torch.float32, device=dev)
    fun_d = torch.empty(m, dtype=torch.float32, device=dev)
    nit_d = torch.empty(m, dtype=torch.int32, device=dev)
    st_d = torch.empty(m, dtype=torch.uint8, device=dev)
    _, fp = t2.preset("gaussian", True)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def problem(n_fit, dense, fused, extras=True):
        p, o = _abi.Problem(), _abi.Outputs()
        keep = _fill_problem(p, "gaussian", fp, te, False, False, 0, 0.0, "loglinear")
        p.echoes, p.memory, p.layout, p.mask_idx = y_d.data_ptr(), _abi.MEM_DEVICE, _abi.LAYOUT_AOS, idx_d.data_ptr()
        p.n_vox, p.n_fit = n_vox, n_fit
        tgt = maps if dense else comp
        o.t2, o.k, o.sigma, o.res = tgt[0].data_ptr(), tgt[1].data_ptr(), tgt[2].data_ptr(), tgt[3].data_ptr()
        if extras:
            o.fun, o.nit, o.status = fun_d.data_ptr(), nit_d.data_ptr(), st_d.data_ptr()
        o.dense = int(dense)
        if fused:
            o.zero_fill_mask = mask_d.data_ptr()
        return p, o, keep

    def runner(p, o):
        def f():
            rc = lib.t2fit_run(C.byref(p), C.byref(o), stream)
            assert rc == 0, lib.t2fit_last_error()
        return f

    res = {}
    if os.environ.get("MB_ONLY") == "variants":     # fill route (T2FIT_FILL is read per call)
        for route in ("stream", "fused"):
            os.environ["T2FIT_FILL"] = route
            p, o, k5 = problem(m, True, True)
            f = runner(p, o)
            print(f"fill {route}: step, windows of 1000 reps:", [round(timeit(f, reps=1000), 1) for _ in range(4)])
        return
    if os.environ.get("MB_ONLY") == "step":         # for ncu: a few whole steps (fit + zero-fill), then a few plain fits
        p, o, k5 = problem(m, True, True)
        f = runner(p, o)
        for _ in range(6):
            f()
        p, o, k4 = problem(m, True, False)
        f = runner(p, o)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        return
    if os.environ.get("MB_ONLY") == "fill":
        p, o, k1 = problem(0, True, True)
        print("fill only", timeit(runner(p, o), reps=10))
        return
    if os.environ.get("MB_ONLY") == "fit":
        p, o, k4 = problem(m, True, False)
        print("fit only", timeit(runner(p, o), reps=10))
        return
    if os.environ.get("MB_ONLY") == "long":
        p, o, k5 = problem(m, True, True)
        f = runner(p, o)
        print("fused, windows of 2000 reps:", [round(timeit(f, reps=2000), 1) for _ in range(8)])
        p, o, k4 = problem(m, True, False)
        f = runner(p, o)
        print("fit only, windows of 2000 reps:", [round(timeit(f, reps=2000), 1) for _ in range(4)])
        print("zero_ only, windows of 2000 reps:", [round(timeit(lambda: maps.zero_(), reps=2000), 1) for _ in range(4)])
        return
    res["torch zero_ 4 maps"] = timeit(lambda: maps.zero_())
    p, o, k1 = problem(0, True, True); res["fill only (fused kernel, n_fit=0)"] = timeit(runner(p, o))
    p, o, k2 = problem(m, False, False); res["fit only, compact out"] = timeit(runner(p, o))
    p, o, k3 = problem(m, False, False, extras=False); res["fit only, compact out, no fun/nit/status"] = timeit(runner(p, o))
    p, o, k4 = problem(m, True, False); res["fit only, dense scatter (maps not zeroed)"] = timeit(runner(p, o))
    p, o, k5 = problem(m, True, True); res["fused fill + fit dense"] = timeit(runner(p, o))
    print("kernel variant:", os.environ.get("T2FIT_KERNEL", "default"))
    for k, v in res.items():
        print(f"  {k:50s} {v:9.2f} us")


if __name__ == "__main__":
    main()
