import os, sys, ctypes
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from cosmomap2_b200 import _device as dv
from torch.multiprocessing.reductions import reduce_tensor
n = 1000
t = torch.full((n,), float(rank + 1), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
gathered = [None] * world
dist.all_gather_object(gathered, (lr, reduce_tensor(t)))
peer = None
for g, (pdev, (fn, args)) in enumerate(gathered):
    if g != rank:
        print(rank, "peer dev", pdev, "can access", torch.cuda.can_device_access_peer(lr, pdev), flush=True)
        dv.call("cm2_enable_peer_access", int(pdev))
        args = list(args); args[6] = lr          # open the IPC mapping in MY device's context
        peer = fn(*args)
        print(rank, "peer tensor device", peer.device, "ptr", hex(peer.data_ptr()), flush=True)
        print(rank, "peer via torch copy", peer[:3].cpu().tolist(), flush=True)
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        dv.call("cm2_dot", peer.data_ptr(), peer.data_ptr(), n, out.data_ptr(), dv.stream())
        torch.cuda.synchronize()
        print(rank, "kernel read of peer memory: dot =", out.item(), "expected", n * float(g + 1) ** 2, flush=True)
dist.barrier()
del peer
dist.barrier()
dist.destroy_process_group()
