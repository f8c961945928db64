"""Throughput on the other BASELINE.json configurations (not the bench line; evidence for DESIGN.md):
   c3  NIST phantom 160x256x256 x 12 TE, gaussian_rician, --no_prior   (both solvers, device-resident)
   c4  batch of fetal-brain volumes 160^3 x 6 TE, gaussian              (t2map_series: staged from host per volume)
   c5  stress 512^3 x 16 TE unmasked, gaussian_rician                   (fast solver on the full size generated on device;
                                                                          faithful solver on a 4 M voxel slab)
   python tools/bench_configs.py [c3] [c4] [c5]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import presets, synth                      # noqa: E402


def timed(fn, reps=3):          # best of `reps`: the first two calls of a process carry one-time module loading
    best = 1e30
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def c3():
    y, mask, te, _ = synth.make_volume("c3", scale=1.0)
    _, fp = presets.preset("gaussian_rician", True)
    dev = torch.device("cuda", 0)
    yt = torch.from_numpy(y.reshape(-1, y.shape[-1])).to(dev)
    idx = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).to(dev)
    for solver in ("fast", "lbfgsb"):
        dt, r = timed(lambda: t2.fit_voxels_batch(yt, idx, te, "gaussian_rician", fp, False, False, solver=solver), reps=4)
        print(f"c3 {tuple(y.shape)} M={idx.numel()} gaussian_rician solver={solver}: {dt*1e3:.1f} ms -> {idx.numel()/dt:.3e} fits/s, "
              f"mean nit {r.nit.float().mean().item():.2f}, failed {(r.status != 0).sum().item()}", flush=True)


def c4(n_vol=16):
    vols = []
    te = None
    for v in range(n_vol):
        y, mask, te, _ = synth.make_volume("c4", scale=1.0, volume_index=v % 4)
        vols.append(([np.ascontiguousarray(y[..., e]) for e in range(y.shape[-1])], [mask.astype(np.uint8)] * y.shape[-1]))
    _, fp = presets.preset("gaussian", True)
    m_tot = sum(int(v[1][0].sum()) for v in vols)
    def consume(depth):                      # as process_t2maps does: use the maps of a volume (here: a checksum), then drop them
        acc = 0.0
        for maps in t2.t2map_series(vols, te, "gaussian", fp, prior=False, depth=depth):
            acc += float(maps.t2[maps.t2.shape[0] // 2].sum())
        return acc
    for depth in (1, 2, 3):
        consume(depth)
        dt, out = timed(lambda: consume(depth), reps=3)
        print(f"c4 {n_vol} volumes 160^3 x {len(te)} TE ({m_tot} masked voxels) t2map_series depth={depth}: {dt*1e3:.1f} ms "
              f"({dt/n_vol*1e3:.2f} ms/volume incl. host cast, H2D of {vols[0][0][0].nbytes*len(te)/1e6:.0f} MB, D2H of 4 maps) "
              f"-> {m_tot/dt:.3e} fits/s end to end", flush=True)


def c5():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(4)
    n, E = 512 ** 3, 16
    te = np.linspace(100, 700, E)
    ted = torch.tensor(te, device=dev, dtype=torch.float32)
    y = torch.empty((n, E), dtype=torch.float32, device=dev)
    step = 1 << 24
    for a in range(0, n, step):                                   # Rician synthetic data generated on the device (8.6 GB)
        b = min(n, a + step)
        t2v = torch.exp(torch.empty(b - a, device=dev).uniform_(np.log(10.0), np.log(2000.0), generator=g))
        s0 = torch.empty(b - a, device=dev).uniform_(300.0, 3000.0, generator=g)
        s = s0[:, None] * torch.exp(-ted[None, :] / t2v[:, None])
        n1 = torch.randn((b - a, E), device=dev, generator=g) * 20.0
        n2 = torch.randn((b - a, E), device=dev, generator=g) * 20.0
        y[a:b] = torch.sqrt((s + n1) ** 2 + n2 ** 2)
        del t2v, s0, s, n1, n2
    _, fp = presets.preset("gaussian_rician", True)
    dt, r = timed(lambda: t2.fit_voxels_batch(y, None, te, "gaussian_rician", fp, False, False, solver="fast"), reps=4)
    print(f"c5 512^3 x 16 TE unmasked M={n} gaussian_rician solver=fast: {dt*1e3:.1f} ms -> {n/dt:.3e} fits/s, "
          f"mean passes {r.nit.float().mean().item():.2f}, failed {(r.status != 0).sum().item()}", flush=True)
    del r
    m = 1 << 22
    dt, r = timed(lambda: t2.fit_voxels_batch(y[:m], None, te, "gaussian_rician", fp, False, False, solver="lbfgsb"), reps=1)
    print(f"c5 slab of {m} voxels gaussian_rician solver=lbfgsb: {dt*1e3:.1f} ms -> {m/dt:.3e} fits/s "
          f"(full volume: {n/(m/dt):.1f} s on one GPU), mean nit {r.nit.float().mean().item():.2f}", flush=True)


if __name__ == "__main__":
    t2.init(0)
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    for w in which:
        {"c3": c3, "c4": c4, "c5": c5}[w]()
