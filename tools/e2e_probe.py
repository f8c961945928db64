"""Where does the host-memory (e2e) path spend its time?  python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["T2FIT_HOST_PROFILE"] = "1"
import torch                                                          # noqa: E402
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import synth                               # noqa: E402

y, mask, te, _ = synth.make_volume("c2", scale=1.0)
flat = np.ascontiguousarray(y.reshape(-1, te.size))
idx = np.flatnonzero(mask.reshape(-1)).astype(np.int64)
_, fp = t2.preset("gaussian", True)
t2.init(0)
print("cpus", len(os.sched_getaffinity(0)), flush=True)
for i in range(6):
    t0 = time.perf_counter()
    r = t2.fit_voxels_batch(flat, idx, te, "gaussian", fp, prior=False)
    t1 = time.perf_counter()
    print(f"call {i}: {1e3*(t1-t0):.3f} ms", flush=True)
# raw copy bandwidths for reference
a = torch.empty(32 << 20, dtype=torch.uint8, pin_memory=True)
d = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(a, non_blocking=True)), ("D2H", lambda: a.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print(name, "32 MiB pinned:", 10 * 32 / 1024 / (time.perf_counter() - t0), "GiB/s")
# host gather alone
t0 = time.perf_counter(); g = flat[idx]; print("numpy gather flat[idx]:", 1e3 * (time.perf_counter() - t0), "ms")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    r = t2.fit_voxels_batch(flat, idx, te, "gaussian", fp, prior=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# page-locked input: no staging (run_host_mapped)
flat_p = t2.pinned_array(None, like=flat)
idx_p = t2.pinned_array(None, like=idx)
for name, a, b in (("pinned input, pageable idx", flat_p, idx), ("pinned input, pinned idx", flat_p, idx_p)):
    ts = []
    for i in range(8):
        t0 = time.perf_counter()
        rp = t2.fit_voxels_batch(a, b, te, "gaussian", fp, prior=False)
        ts.append(1e3 * (time.perf_counter() - t0))
    print(name, ["%.3f" % x for x in ts], flush=True)
    assert np.array_equal(rp.t2, r.t2) and np.array_equal(rp.status, r.status)
