#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/test_ground_dist.py : the ground-template filter on a TOD sharded
by detector equals the unsharded one (development check, N GPUs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import distributed
    ndet, ns = 8, 50001
    rng = np.random.default_rng(5)
    ground = np.concatenate([((np.arange(ns) * 0.01 + 3 * b) % 150).astype(np.int64) for b in range(ndet)])
    ground[rng.random(ndet * ns) < 0.03] = -1
    v = rng.standard_normal(ndet * ns)
    ref = cm.GroundFilterLO(ground.copy()) * v                    # unsharded, on every rank
    lo, hi = distributed.shard_detectors(ndet, world, rank)
    sl = slice(lo * ns, hi * ns)
    if rank == 1:                                                 # a rank that does not see the highest bins
        g_loc = np.where(ground[sl] > 100, -1, ground[sl])
        ref_mod = ground.copy()
        ref_mod[sl] = g_loc
    else:
        g_loc = ground[sl].copy()
        ref_mod = None
    # all ranks must agree on the modified reference: rank 1's modification is broadcast
    obj = [ref_mod]
    dist.broadcast_object_list(obj, src=1 if world > 1 else 0)
    if obj[0] is not None:
        ref = cm.GroundFilterLO(obj[0].copy()) * v
    G = cm.GroundFilterLO(g_loc, comm=True)
    out = G * v[sl]
    err = np.max(np.abs(out - ref[sl])) / np.max(np.abs(ref))
    t = torch.tensor([err], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("ground filter sharded over %d GPUs: max rel err %.2e, nbins %d" % (world, t.item(), G.nbins))
    assert t.item() < 1e-12
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
