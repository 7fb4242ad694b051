#!/usr/bin/env python
"""
BASELINE.json configs[2] end to end: correlated noise + subscan filtering, M_BD-preconditioned CG.

    A = P^T F N^-1 F P ,   b = P^T F N^-1 F d

with N^-1 = BlockLO(ns, bands, offdiag=True) -- one symmetric banded Toeplitz block per detector,
4096 coefficients by default (the FFT overlap-save kernel) -- F = FilterLO(...) the subscan offset
filter, and M_BD built with the per-detector weights a_0 (the reference feeds `N.diag` to
ProcessTimeSamples, src/test_BD_precond_onto_real_data.py:78-80).

    python examples/solve_correlated.py --nt 1.25e8 --ndet 8                    # one GPU's share
    torchrun --nproc-per-node 8 examples/solve_correlated.py --nt 1.25e8 --ndet 8   # 1e9 samples, 64 detectors

The pointing is generated on the device (inputs only).  Under torchrun the TOD is sharded by detector:
noise blocks and subscans never straddle detectors, so the time domain needs no exchange and every
A apply ends in one sum of the map-domain vector.  The work itself lives in cosmomap2_b200/workloads.py
(bench.py runs the same function for its `secondary` block); see `plan` in the output for the kernels
the A apply runs as.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=float, default=1.25e8, help="samples per GPU")
    ap.add_argument("--ndet", type=int, default=8, help="detectors (noise blocks) per GPU")
    ap.add_argument("--nband", type=int, default=4096, help="Toeplitz coefficients per detector")
    ap.add_argument("--nside", type=int, default=512)
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--ny", type=int, default=500)
    ap.add_argument("--rtol", type=float, default=1e-6)
    ap.add_argument("--maxiter", type=int, default=300)
    ap.add_argument("--time-iters", type=int, default=20, help="A applies timed with CUDA events after the solve")
    ap.add_argument("--two-level", type=int, default=0, metavar="R",
                    help="also solve with M_2lvl on an R-dimensional scan coarse space (cosmomap2_b200.scan_coarse_space)")
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from cosmomap2_b200 import workloads
    out = workloads.correlated(nt=args.nt, ndet=args.ndet, nband=args.nband, nside=args.nside, nx=args.nx, ny=args.ny,
                               rtol=args.rtol, maxiter=args.maxiter, time_iters=args.time_iters, two_level_r=args.two_level)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
