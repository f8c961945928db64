"""Seeded synthetic multi-echo volumes for the five BASELINE.json configurations.

There is no patient data (and no network); tests and ``bench.py`` use these
generators.  Shapes, TEs, masks, tissue values and seeds follow SURVEY.md §8(d).
Every generator returns ``(t2w f32[z,y,x,E], mask bool[z,y,x], TEeffs f64[E], truth)``
in the layout ``process_t2maps`` builds (run_t2mapping.py:383-386): echoes stacked on
the last axis, C order.
"""
from __future__ import annotations

import numpy as np

__all__ = ["CONFIGS", "make_volume", "ellipsoid_mask", "decay_signal"]

CONFIGS = {
    # name: shape(z,y,x), TEs, fit, field, prior, noise kind
    "c1": dict(shape=(64, 64, 64), te=[114, 150, 202, 299], fit="gaussian", field="lf",
               prior=False, seed=0),
    "c2": dict(shape=(256, 256, 256), te=[114, 132, 150, 176, 202], fit="gaussian", field="lf",
               prior=False, seed=1),
    "c3": dict(shape=(160, 256, 256), te=[113, 133, 152, 176, 203, 230, 254, 296, 340, 400, 480, 600],
               fit="gaussian_rician", field="lf", prior=False, seed=2),
    "c4": dict(shape=(160, 160, 160), te=[114, 132, 150, 176, 202, 229], fit="gaussian", field="lf",
               prior=False, seed=3, volumes=64),
    "c5": dict(shape=(512, 512, 512), te=list(np.linspace(100, 700, 16)), fit="gaussian_rician",
               field="lf", prior=False, seed=4),
}

# NMR ground-truth T2 of the 14 high-field phantom spheres (run_t2mapping.py:24)
_PHANTOM_T2 = [1044, 624, 428, 258, 186, 137, 90, 63, 44, 27, 19, 15, 10, 8]


def ellipsoid_mask(shape, semi_axes, center=None):
    z, y, x = np.ogrid[:shape[0], :shape[1], :shape[2]]
    c = [(s - 1) / 2.0 for s in shape] if center is None else center
    a = semi_axes
    return ((z - c[0]) / a[0]) ** 2 + ((y - c[1]) / a[1]) ** 2 + ((x - c[2]) / a[2]) ** 2 <= 1.0


def decay_signal(s0, t2, te, rng, sigma, rician):
    """S0*exp(-TE/T2) + noise; Gaussian, or Rician sqrt((s+n1)^2+n2^2)."""
    te = np.asarray(te, np.float32)
    s = s0[..., None].astype(np.float32) * np.exp(-te / t2[..., None].astype(np.float32))
    n1 = rng.standard_normal(s.shape, dtype=np.float32) * np.float32(sigma)
    if not rician:
        return (s + n1).astype(np.float32)
    n2 = rng.standard_normal(s.shape, dtype=np.float32) * np.float32(sigma)
    return np.sqrt((s + n1) ** 2 + n2 ** 2).astype(np.float32)


def _scaled_shape(shape, scale):
    return tuple(max(8, int(round(s * scale))) for s in shape)


def make_volume(name: str, scale: float = 1.0, volume_index: int = 0):
    """Build configuration ``name`` ('c1'..'c5'); ``scale`` shrinks every spatial axis
    (tests use small scales; the bench uses 1.0)."""
    cfg = CONFIGS[name]
    shape = _scaled_shape(cfg["shape"], scale)
    te = np.asarray(cfg["te"], np.float64)
    rng = np.random.default_rng([cfg["seed"], volume_index])
    n = shape[0] * shape[1] * shape[2]
    if name == "c1":
        mask = ellipsoid_mask(shape, [0.4 * s for s in shape])
        t2 = rng.uniform(60, 300, shape).astype(np.float32)
        s0 = rng.uniform(200, 900, shape).astype(np.float32)
        y = decay_signal(s0, t2, te, rng, 8.0, rician=False)
    elif name == "c2":
        ax = [65 / 256 * shape[0], 85 / 256 * shape[1], 70 / 256 * shape[2]]
        mask = ellipsoid_mask(shape, ax)
        wm = ellipsoid_mask(shape, [a * 0.72 for a in ax])
        csf = ellipsoid_mask(shape, [a * 0.16 for a in ax])
        t2 = rng.normal(165, 35, shape).astype(np.float32)                 # GM shell
        t2 = np.where(wm, rng.normal(118, 15, shape).astype(np.float32), t2)
        t2 = np.where(csf, rng.uniform(600, 2000, shape).astype(np.float32), t2)
        # sparse sulcal CSF blobs in the GM shell
        blobs = (rng.random(shape, dtype=np.float32) < 0.01) & ~wm
        t2 = np.where(blobs, rng.uniform(600, 2000, shape).astype(np.float32), t2)
        t2 = np.clip(t2, 20, 2500).astype(np.float32)
        s0 = rng.uniform(250, 700, shape).astype(np.float32)
        y = decay_signal(s0, t2, te, rng, 12.0, rician=False)
    elif name == "c3":
        r = 100 / 256 * shape[1]
        mask = ellipsoid_mask(shape, [min(r, 0.49 * shape[0]), r, r])
        t2 = np.exp(rng.uniform(np.log(10), np.log(2000), shape)).astype(np.float32)
        # 14 ROI spheres on a ring in the central slab
        cz, cy, cx = [(s - 1) / 2.0 for s in shape]
        for i, tv in enumerate(_PHANTOM_T2):
            ang = 2 * np.pi * i / len(_PHANTOM_T2)
            cen = [cz, cy + 0.6 * r * np.sin(ang), cx + 0.6 * r * np.cos(ang)]
            roi = ellipsoid_mask(shape, [0.08 * r * 1.5] * 3, center=cen)
            t2 = np.where(roi, np.float32(tv), t2)
        s0 = rng.uniform(600, 3000, shape).astype(np.float32)
        y = decay_signal(s0, t2, te, rng, 20.0, rician=True)
    elif name == "c4":
        ax = [0.26 * shape[0], 0.32 * shape[1], 0.27 * shape[2]]
        mask = ellipsoid_mask(shape, ax)
        t2 = rng.uniform(80, 400, shape).astype(np.float32)
        s0 = rng.uniform(200, 900, shape).astype(np.float32)
        y = decay_signal(s0, t2, te, rng, 12.0, rician=False)
    elif name == "c5":
        mask = np.ones(shape, bool)
        t2 = np.exp(rng.uniform(np.log(10), np.log(2000), shape)).astype(np.float32)
        s0 = rng.uniform(300, 3000, shape).astype(np.float32)
        y = decay_signal(s0, t2, te, rng, 20.0, rician=True)
    else:
        raise KeyError(name)
    assert y.shape == shape + (te.size,) and y.size == n * te.size
    y = np.where(mask[..., None], y, np.abs(y) * 0.02).astype(np.float32)   # faint background
    return y, mask, te, {"t2": t2, "s0": s0, "cfg": dict(cfg, shape=shape)}
