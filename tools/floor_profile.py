"""One launch pair of the 3-parameter fast solver on a c5-like slab (unmasked, 16 TE, Rician data generated on the
device) for ncu:  python tools/floor_profile.py [n_voxels]   (T2FIT_FLOOR_KERNEL=oneshot|queue selects the kernel)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import presets                             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 23
E = int(os.environ.get("FP_E", "16"))
t2.init(0)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(4)
te = np.linspace(100, 700, E) if E > 3 else np.array([114.0, 202.0, 299.0])[:E]
ted = torch.tensor(te, device=dev, dtype=torch.float32)
t2v = torch.exp(torch.empty(n, device=dev).uniform_(np.log(10.0), np.log(2000.0), generator=g))
s0 = torch.empty(n, device=dev).uniform_(300.0, 3000.0, generator=g)
s = s0[:, None] * torch.exp(-ted[None, :] / t2v[:, None])
y = torch.sqrt((s + torch.randn((n, E), device=dev, generator=g) * 20.0) ** 2 + (torch.randn((n, E), device=dev, generator=g) * 20.0) ** 2)
del s, t2v, s0
_, fp = presets.preset("gaussian_rician", True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
for _ in range(int(os.environ.get("FP_REPS", "2"))):
    ev[0].record()
    r = t2.fit_voxels_batch(y, None, te, "gaussian_rician", fp, False, False, solver="fast")
    ev[1].record()
    torch.cuda.synchronize()
    times.append(round(ev[0].elapsed_time(ev[1]), 2))
print(os.environ.get("T2FIT_FLOOR_KERNEL", "queue"), "refill", os.environ.get("T2FIT_QUEUE_REFILL", "default"), "E", E, n, f"{ev[0].elapsed_time(ev[1]):.2f} ms", times, "mean passes", r.nit.float().mean().item())
