"""A/B of the reference-faithful (L-BFGS-B) kernels, device-resident: thread-per-voxel vs cooperative (8 / 16 / 32 lanes per
voxel) vs the dense-matrix form (dense).  python tools/lb_bench.py [c2] [c3] [c3r] [c5] [--kernels thread,coop8,...] [--scale S]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import presets, synth                      # noqa: E402


def workload(name, scale):
    if name == "c5":
        dev = torch.device("cuda", 0)
        g = torch.Generator(device=dev).manual_seed(4)
        n, E = int((1 << 21) * scale), 16
        te = np.linspace(100, 700, E)
        ted = torch.tensor(te, device=dev, dtype=torch.float32)
        t2v = torch.exp(torch.empty(n, device=dev).uniform_(np.log(10.0), np.log(2000.0), generator=g))
        s0 = torch.empty(n, device=dev).uniform_(300.0, 3000.0, generator=g)
        s = s0[:, None] * torch.exp(-ted[None, :] / t2v[:, None])
        y = torch.sqrt((s + torch.randn((n, E), device=dev, generator=g) * 20.0) ** 2 + (torch.randn((n, E), device=dev, generator=g) * 20.0) ** 2)
        return y, None, te, "gaussian_rician"
    cfg = {"c2": "c2", "c3": "c3", "c3r": "c3"}[name]
    y, mask, te, _ = synth.make_volume(cfg, scale=scale)
    yt = torch.from_numpy(np.ascontiguousarray(y.reshape(-1, y.shape[-1]))).cuda()
    idx = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).cuda()
    return yt, idx, te, {"c2": "gaussian", "c3": "gaussian_rician", "c3r": "rician"}[name]


def main():
    args = sys.argv[1:]
    kernels = ["thread", "coop8", "coop16", "coop32"]
    scale = None
    names = []
    i = 0
    while i < len(args):
        if args[i] == "--kernels":
            kernels = args[i + 1].split(","); i += 2
        elif args[i] == "--scale":
            scale = float(args[i + 1]); i += 2
        else:
            names.append(args[i]); i += 1
    t2.init(0)
    for name in names or ["c2", "c3"]:
        sc = scale if scale is not None else {"c2": 0.6, "c3": 0.5, "c3r": 0.5, "c5": 1.0}[name]
        y, idx, te, fit = workload(name, sc)
        _, fp = presets.preset(fit, True)
        m = y.shape[0] if idx is None else idx.numel()
        ref = None
        for k in kernels:
            os.environ["T2FIT_LB_KERNEL"] = k if k != "dense" else "thread"
            solver = "lbfgsb_dense" if k == "dense" else "lbfgsb"
            best = 1e30
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r = t2.fit_voxels_batch(y, idx, te, fit, fp, False, False, solver=solver, check_bounds=False)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            same = ""
            if ref is None:
                ref = r
            else:
                rel = ((r.t2 - ref.t2).abs() / ref.t2.abs())
                same = (f", identical to {kernels[0]}: {float(((r.t2 == ref.t2) & (r.nit == ref.nit)).float().mean()):.5f}, T2 within 1e-3: "
                        f"{float((rel <= 1e-3).float().mean()):.5f}, nit equal {float((r.nit == ref.nit).float().mean()):.5f}")
            print(f"{name} scale {sc} M={m} E={len(te)} {fit} kernel={k}: {best*1e3:.1f} ms -> {m/best:.3e} fits/s, mean nit "
                  f"{r.nit.float().mean().item():.2f}, failed {(r.status != 0).sum().item()}{same}", flush=True)
        os.environ.pop("T2FIT_LB_KERNEL", None)


if __name__ == "__main__":
    main()
