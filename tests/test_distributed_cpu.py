"""N > 1 host logic on CPU: world_size-2 gloo processes partition the masked list into slabs, fit their
slab (stand-in fit function: the host simulation -- test tool, the product fit is CUDA) and all-gather."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fetal_t2mapping_b200 import distributed as D
from tests.conftest import ROOT


def test_slab_bounds_properties():
    for n in (0, 1, 127, 128, 129, 1000, 1619960, 134217728):
        for w in (1, 2, 3, 4, 8):
            b = D.slab_bounds(n, w)
            L = D.slab_length(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))          # contiguous, ordered
            assert L % 128 == 0 and all(x == min(n, r * L) for r, (x, _) in enumerate(b))   # slab r starts at r L: the
            sizes = [y - x for x, y in b]                                     # L-padded slabs concatenate to the full vector
            assert min(sizes) >= 0 and max(sizes) <= L and sum(sizes) == n
            assert n == 0 or L - -(-n // w) < 128                              # balanced to within one alignment unit


class _R:
    pass


def _hostsim_fit(rows, idx, te, fit, fp, prior, norm, **kw):
    from tests import hostsim
    o = hostsim.fit(rows[idx], te, fit, fp["initial_guess"], fp["param_bounds"], prior, norm)
    r = _R()
    r.t2, r.k, r.sigma, r.res, r.status = o["t2"], o["k"], o["sigma"], o["res"], o["status"]
    return r


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fetal_t2mapping_b200 import presets, synth
    y, mask, te, _ = synth.make_volume("c1", scale=0.3)
    flat = y.reshape(-1, te.size)
    idx = np.flatnonzero(mask.reshape(-1))
    _, fp = presets.preset("gaussian", True)
    out = D.fit_voxels_sharded(flat, idx, te, "gaussian", fp, prior=False, fit_fn=_hostsim_fit)
    np.save(os.path.join(tmp, f"t2_{rank}.npy"), out["t2"].numpy())
    np.save(os.path.join(tmp, f"st_{rank}.npy"), out["status"].numpy())
    assert out["status"].dtype == torch.uint8
    # the same job with every rank holding ONLY the rows of its own slab
    a, b = D.slab_bounds(idx.size, world)[rank]
    out2 = D.fit_slab_sharded(np.ascontiguousarray(flat[idx[a:b]]), idx.size, te, "gaussian", fp, prior=False,
                              fit_fn=lambda rows, _i, *a_, **k_: _hostsim_fit(rows, np.arange(rows.shape[0]), *a_, **k_))
    assert torch.equal(out2["t2"], out["t2"]) and torch.equal(out2["res"], out["res"])
    # a fit that raises on ONE rank only (scipy's ValueError under --no_prior): every rank raises, nobody hangs in the gather
    def raising(rows, i, *a_, **k_):
        if rank == 1:
            raise ValueError("An upper bound is less than the corresponding lower bound.")
        return _hostsim_fit(rows, i, *a_, **k_)
    try:
        D.fit_voxels_sharded(flat, idx, te, "gaussian", fp, prior=False, fit_fn=raising)
        raised = "no"
    except ValueError as e:
        raised = str(e)
    open(os.path.join(tmp, f"raised_{rank}.txt"), "w").write(raised)
    if rank == 0:
        single = _hostsim_fit(flat, idx, te, "gaussian", fp, False, False)
        np.save(os.path.join(tmp, "single.npy"), single.t2)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_fit_equals_single(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npy")
    for r in range(world):
        got = np.load(tmp_path / f"t2_{r}.npy")
        assert got.shape == single.shape and np.array_equal(got, single)      # every rank holds the full vector
        assert (np.load(tmp_path / f"st_{r}.npy") == 0).all()
        assert "upper bound" in open(tmp_path / f"raised_{r}.txt").read()
