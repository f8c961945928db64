"""GPU check of the reference-faithful solver (T2FIT_SOLVER_LBFGSB): agreement with the golden
fixtures (reference outputs) and throughput.  Run under gpurun:  python tools/lb_check.py"""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import presets, synth                      # noqa: E402


def fixture_agreement():
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "c*_*.npz"))):
        d = np.load(path, allow_pickle=True)
        fit, field, prior = str(d["fit"]), str(d["field"]), bool(d["prior"])
        _, fp = presets.preset(fit, field == "lf")
        r = t2.fit_voxels_batch(d["rows"], None, d["te"], fit, fp, prior, False, solver="lbfgsb")
        ref = d["ref_params"]
        rel = np.abs(r.t2.astype(np.float64) - ref[:, 1]) / np.abs(ref[:, 1])
        print(f"{os.path.basename(path):28s} M={ref.shape[0]:5d} T2 rel<=1e-3 {np.mean(rel <= 1e-3):.4f} <=1e-5 {np.mean(rel <= 1e-5):.4f} "
              f"nit eq {np.mean(r.nit == d['ref_nit']):.4f} success eq {np.mean((r.status == 0) == d['ref_success']):.4f}", flush=True)


def throughput(cfg, scale, fit=None, reps=3):
    y, mask, te, _ = synth.make_volume(cfg, scale=scale)
    c = synth.CONFIGS[cfg]
    fit = fit or c["fit"]
    _, fp = presets.preset(fit, c["field"] == "lf")
    dev = torch.device("cuda", 0)
    yt = torch.from_numpy(y.reshape(-1, y.shape[-1])).to(dev)
    idx = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).to(dev)
    for solver in ("lbfgsb", "fast"):
        if solver == "fast" and fit == "rician":
            continue
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = t2.fit_voxels_batch(yt, idx, te, fit, fp, c["prior"], False, solver=solver)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print(f"{cfg} scale {scale} {fit} E={len(te)} M={idx.numel()} solver={solver}: best {min(ts)*1e3:.2f} ms "
              f"-> {idx.numel()/min(ts):.3e} fits/s; mean nit {r.nit.float().mean().item():.2f}", flush=True)


if __name__ == "__main__":
    t2.init(0)
    fixture_agreement()
    throughput("c2", 1.0)
    throughput("c3", 0.5)
    throughput("c3", 0.5, fit="rician")
