"""N > 1 on real GPUs (skipped unless the box shows at least two): one process per GPU, NCCL; every rank fits its slab of
the masked list with the CUDA path and ONE all-gather stitches the parameter vectors (distributed.fit_voxels_sharded)."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, fit):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fetal_t2mapping_b200 as t2
    from fetal_t2mapping_b200 import distributed as D, synth
    t2.init(rank)
    y, mask, te, _ = synth.make_volume("c1", scale=0.35)
    flat = torch.from_numpy(np.abs(y.reshape(-1, te.size)) + 1.0).cuda()
    idx = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).cuda()
    _, fp = t2.preset(fit, True)
    out = D.fit_voxels_sharded(flat, idx, te, fit, fp, prior=False)
    torch.cuda.synchronize()
    fused = D.fit_voxels_fused_gather(flat, idx, te, fit, fp, prior=False, root=0)     # kernels store into rank 0's buffer
    assert (fused is None) == (rank != 0)
    if rank == 0:
        for name in ("t2", "k", "sigma", "res"):
            assert torch.equal(fused[name], out[name]), name
        assert torch.equal(fused["status"], out["status"])
    # every rank holding ONLY the rows of its own slab (what a loader hands each GPU)
    a, b = D.slab_bounds(idx.numel(), world)[rank]
    out2 = D.fit_slab_sharded(flat[idx[a:b]].contiguous(), idx.numel(), te, fit, fp, prior=False)
    for name in ("t2", "k", "sigma", "res", "status"):
        assert torch.equal(out2[name], out[name]), name
    assert out["status"].dtype == torch.uint8
    # one rank's slab holds a voxel scipy rejects (--no_prior, S(TE0) > 10000): EVERY rank raises, nobody hangs in the gather
    bad = flat.clone()
    bad[idx[idx.numel() - 3], 0] = 2.0e4                       # lies in the last rank's slab
    for call in (lambda: D.fit_voxels_sharded(bad, idx, te, fit, fp, prior=False),
                 lambda: D.fit_voxels_fused_gather(bad, idx, te, fit, fp, prior=False, root=0)):
        try:
            call()
            raised = False
        except ValueError as e:
            raised = "upper bound" in str(e)
        assert raised, f"rank {rank} did not raise"
    # a stream of jobs through the double-buffered pipeline (gather of job i overlaps the fit of job i+1)
    pipe = D.SlabPipeline(idx.numel(), fit)
    rows = flat[idx[a:b]].contiguous()
    rows2 = (rows * 1.5).contiguous()
    slots = [pipe.submit(r_, te, fp, prior=False) for r_ in (rows, rows2, rows)]
    last = pipe.result(slots[2], check=True)
    torch.cuda.synchronize()
    for name in ("t2", "k", "res", "status"):
        assert torch.equal(last[name], out[name]), name
    mid = pipe.result(slots[1])
    torch.cuda.synchronize()
    single_mid = t2.fit_voxels_batch((flat * 1.5).contiguous(), idx, te, fit, fp, prior=False)
    assert torch.equal(mid["t2"], single_mid.t2) and torch.equal(mid["k"], single_mid.k) and not torch.equal(mid["k"], out["k"])
    pipe.drain()
    # the all-gather fused into the kernels (peer stores into every rank's buffer + a one-element all-reduce as the barrier)
    fused_ag = D.FusedAllGather(idx.numel(), fit)
    s0 = fused_ag.submit(rows, te, fp, prior=False)
    s1 = fused_ag.submit(rows2, te, fp, prior=False)
    r0, r1 = fused_ag.result(s0, check=True), fused_ag.result(s1)
    torch.cuda.synchronize()
    for name in ("t2", "k", "res", "status") + (() if fit == "gaussian" else ("sigma",)):
        assert torch.equal(r0[name], out[name]), name
    assert torch.equal(r1["t2"], single_mid.t2) and torch.equal(r1["k"], single_mid.k)
    s2 = fused_ag.submit(rows, te, fp, prior=False)                    # slot 0 again
    assert s2 == s0
    torch.cuda.synchronize()
    assert torch.equal(fused_ag.result(s2)["res"], out["res"])
    fused_ag.close()
    # the same through the NVSwitch multicast address of the buffers (one store per value, replicated by the switch); on a
    # node without multicast the constructor agrees on the unicast form on every rank and the check is the one above again
    fused_mc = D.FusedAllGather(idx.numel(), fit, multicast="on")
    slots_mc = [fused_mc.submit(r_, te, fp, prior=False) for r_ in (rows, rows2, rows)]
    rm2, rm1 = fused_mc.result(slots_mc[2], check=True), fused_mc.result(slots_mc[1])
    torch.cuda.synchronize()
    for name in ("t2", "k", "res", "status") + (() if fit == "gaussian" else ("sigma",)):
        assert torch.equal(rm2[name], out[name]), ("multicast", name)
    assert torch.equal(rm1["t2"], single_mid.t2) and torch.equal(rm1["k"], single_mid.k)
    print(f"rank {rank}: fused all-gather multicast = {fused_mc.multicast}", flush=True)
    with open(os.path.join(tmp, f"mc_{rank}.txt"), "w") as fh:
        fh.write(str(int(fused_mc.multicast)))
    fused_mc.close()
    again = D.fit_voxels_sharded(flat, idx, te, fit, fp, prior=False)      # and the next clean job is unaffected
    assert torch.equal(again["t2"], out["t2"])
    np.save(os.path.join(tmp, f"t2_{rank}.npy"), out["t2"].cpu().numpy())
    np.save(os.path.join(tmp, f"st_{rank}.npy"), out["status"].cpu().numpy())
    if rank == 0:
        single = t2.fit_voxels_batch(flat, idx, te, fit, fp, prior=False)
        torch.cuda.synchronize()
        np.save(os.path.join(tmp, "single.npy"), single.t2.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fit", ["gaussian", "gaussian_rician"])
def test_two_gpu_sharded_fit_equals_single_gpu(tmp_path, fit):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path), fit), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npy")
    for r in range(world):
        got = np.load(tmp_path / f"t2_{r}.npy")
        assert got.shape == single.shape and np.array_equal(got, single)       # every rank holds the full, identical vector
        assert (np.load(tmp_path / f"st_{r}.npy") == 0).all()
