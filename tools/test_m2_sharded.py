#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/test_m2_sharded.py : the pixel-sharded two-level preconditioner
equals the replicated one and is timed beside it (development check, N GPUs)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import distributed, synthetic
    pol, r = 3, 32
    sc = synthetic.raster_scan(20000000, nside=512, ndet=16, nx=1001, ny=499, samples_per_pixel=8.0, seed=0,
                               with_data=False)
    N = cm.BlockLO(sc.ns, sc.weights)
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    n = pol * npix
    P = cm.SparseLO(npix, sc.nt, sc.pix, pol=pol, angle_processed=pts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A = P.T * N * P
    g = torch.Generator(device="cuda")
    g.manual_seed(3)                                        # the same Z on every rank
    Zt = torch.randn((r, n), dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n)
    AZt = torch.stack([A._apply(Zt[i]) for i in range(r)])
    Zd, AZd = cm.DeflationLO(Zt.t()), cm.DeflationLO(AZt.t())
    E = cm.CoarseLO(Zt.t(), AZt.t(), r, apply="eig")
    M2 = cm.TwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
    M2s = distributed.ShardedTwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
    v = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y0, y1 = M2._apply(v), M2s._apply(v)
    err = float((y1 - y0).abs().max() / y0.abs().max())
    # bit-identical on every rank (replicated PCG vectors must stay replicated)
    same = True
    if world > 1:
        ys = [torch.empty_like(y1) for _ in range(world)]
        dist.all_gather(ys, y1)
        same = all(torch.equal(ys[0], t) for t in ys)

    def timeit(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t0, t1 = timeit(lambda: M2._apply(v)), timeit(lambda: M2s._apply(v))
    if rank == 0:
        print(json.dumps({"world": world, "n": n, "r": r, "npix": int(npix), "rel_err": err, "identical_on_ranks": same,
                          "replicated_ms": t0, "sharded_ms": t1}))
    assert err < 1e-11 and same
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
