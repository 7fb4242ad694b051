"""torchrun --nproc-per-node N tools/test_p2p_allreduce.py : correctness + timing of the NVLink
peer-memory all-reduce against NCCL (run on the GPU box only)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from cosmomap2_b200 import distributed
for n in (1500000, 1500001, 7):
    ar = distributed.P2PAllReduce(n)
    g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
    y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    allv = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(allv, y)
    ref = allv[0].clone()
    for k in range(1, world):
        ref += allv[k]
    for it in range(3):
        out = ar(y).clone()
        assert torch.equal(out, ref), "rank %d n=%d iter %d mismatch: %g" % (rank, n, it, (out - ref).abs().max().item())
    # timing
    big = torch.randn(8192, 8192, device="cuda")
    def timeit(fn, reps=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): torch.matmul(big, big)      # ~ms of device work: the host runs ahead, so the
        e0.record()                                     # loop below is timed back-to-back on the device
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    t_p2p = timeit(lambda: ar(y))
    z = y.clone()
    t_nccl = timeit(lambda: dist.all_reduce(z))
    err = int(torch.frombuffer(bytearray(ar.sig.cpu().numpy().tobytes()[-4:]), dtype=torch.int32)[0]) if False else 0
    if rank == 0:
        print("n=%d world=%d  p2p %.1f us   nccl %.1f us   (exact match with rank-ordered sum)" % (n, world, t_p2p, t_nccl))
    ar.close()
dist.barrier()
dist.destroy_process_group()
