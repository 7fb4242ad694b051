#!/usr/bin/env python
"""
End-to-end two-level solve at scale (the shape of BASELINE.json configs[3]): offset-filtered
map-making  (P^T F P) x = P^T F d  on a synthetic raster scan, first with M_BD, then with the
two-level preconditioner M_2lvl = Mbd*(I - AZd*E*Zd.T) + Zd*E*Zd.T of the reference's
src/test_M2_precond_onto_real_data.py:54-122.  The deflation space is either the a-priori subdomain
space built from the scan order (--coarse scan, the default: no Krylov phase) or the reference's
recipe (--coarse ritz: run_krypy_arnoldi -> find_ritz_eigenvalues -> CoarseLO(apply='eig')).

    python examples/solve_two_level.py --nt 5e8                       # one GPU
    torchrun --nproc-per-node 8 examples/solve_two_level.py --nt 5e8  # 4e9 samples on 8 GPUs

The pointing is generated on the device (inputs only).  With torchrun the TOD is sharded by
detector; the map-domain sums go through cosmomap2_b200.distributed.  The work itself lives in
cosmomap2_b200/workloads.py (bench.py runs the same function for its `secondary` block).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=float, default=2e8, help="samples per GPU")
    ap.add_argument("--nside", type=int, default=1024)
    ap.add_argument("--nx", type=int, default=1600)
    ap.add_argument("--ny", type=int, default=800)
    ap.add_argument("--ndet", type=int, default=64)
    ap.add_argument("--r", "--ncoarse", dest="r", type=int, default=32,
                    help="dimension of the deflation space (under torchrun use --ncoarse: its parser claims --r)")
    ap.add_argument("--coarse", default="scan", choices=["scan", "ritz"],
                    help="deflation space: 'scan' = a-priori subdomain space from the scan order, "
                         "'ritz' = preconditioned Arnoldi + Ritz vectors (the reference's recipe)")
    ap.add_argument("--smooth", type=int, default=2, help="coordinate smoothing sweeps of the scan coarse space")
    ap.add_argument("--arnoldi", type=int, default=300)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--maxiter", type=int, default=2000)
    ap.add_argument("--shard-m2", action="store_true",
                    help="N > 1: Z, AZ and the M_BD blocks sharded by pixel over the ranks "
                         "(distributed.ShardedTwoLevelPreconditionerLO)")
    ap.add_argument("--poly-order", type=int, default=0,
                    help="subscan filter: 0 = offsets (src/test_M2_precond_onto_real_data.py:79-80), "
                         ">0 = Legendre polynomials up to that order (:37-38)")
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from cosmomap2_b200 import workloads
    out = workloads.two_level(nt=args.nt, nside=args.nside, nx=args.nx, ny=args.ny, ndet=args.ndet, r=args.r,
                              coarse=args.coarse, smooth=args.smooth, arnoldi=args.arnoldi, rtol=args.rtol,
                              maxiter=args.maxiter, shard_m2=args.shard_m2, poly_order=args.poly_order)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
