"""The page-locked-input host path (run_host_mapped: one kernel gathers the masked rows from host memory over PCIe) for
ncu:  T2FIT_HOST_IN=mapped python tools/mapped_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("T2FIT_HOST_IN", "mapped")
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import synth                               # noqa: E402

y, mask, te, _ = synth.make_volume("c2", scale=1.0)
flat = np.ascontiguousarray(y.reshape(-1, te.size))
idx = np.flatnonzero(mask.reshape(-1)).astype(np.int64)
_, fp = t2.preset("gaussian", True)
t2.init(0)
flat_p, idx_p = t2.pinned_array(None, like=flat), t2.pinned_array(None, like=idx)
ts = []
for i in range(int(os.environ.get("MP_CALLS", "6"))):
    t0 = time.perf_counter()
    r = t2.fit_voxels_batch(flat_p, idx_p, te, "gaussian", fp, prior=False)
    ts.append(1e3 * (time.perf_counter() - t0))
print(os.environ["T2FIT_HOST_IN"], ["%.3f" % x for x in ts], "ms per call", flush=True)
if os.environ.get("MP_PROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        r = t2.fit_voxels_batch(flat_p, idx_p, te, "gaussian", fp, prior=False)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
