"""One launch of the L-BFGS-B kernel for ncu:  python tools/lb_profile.py [cfg] [scale] [fit] [solver]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fetal_t2mapping_b200 as t2                                    # noqa: E402
from fetal_t2mapping_b200 import presets, synth                      # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
c = synth.CONFIGS[cfg]
fit = sys.argv[3] if len(sys.argv) > 3 else c["fit"]
solver = sys.argv[4] if len(sys.argv) > 4 else "lbfgsb"
y, mask, te, _ = synth.make_volume(cfg, scale=scale)
_, fp = presets.preset(fit, c["field"] == "lf")
t2.init(0)
dev = torch.device("cuda", 0)
yt = torch.from_numpy(y.reshape(-1, y.shape[-1])).to(dev)
idx = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).to(dev)
for _ in range(2):
    r = t2.fit_voxels_batch(yt, idx, te, fit, fp, c["prior"], False, solver=solver)
torch.cuda.synchronize()
print(cfg, scale, fit, idx.numel(), "mean nit", r.nit.float().mean().item())
