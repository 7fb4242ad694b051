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
A apply ends in one sum of the map-domain vector (cosmomap2_b200.distributed.AllReduceLO).
The A apply runs as: subscan means from the run table + one gather pass (F P fused), the Toeplitz
kernel, the subscan filter, the scatter -- see `plan` in the output.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from solve_two_level import make_scan  # noqa: E402


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
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import distributed, synthetic

    pol = 3
    nt, ns, pix, phi, sub_len, sub_start, g = make_scan(int(args.nt), args.nside, args.nx, args.ny, args.ndet, 8.0,
                                                        seed=rank)
    npix_full = 12 * args.nside ** 2
    bands = synthetic.toeplitz_bands(args.ndet, args.nband, seed=100 + rank)
    N = cm.BlockLO(ns, bands, offdiag=True)
    Nw = cm.BlockLO(ns, [a[0] for a in bands])             # the diagonal of N^-1: the weights of M_BD
    pts = cm.ProcessTimeSamples(pix, npix_full, obspix=np.arange(npix_full), pol=pol, phi=phi, w=Nw.diag,
                                comm=(True if world > 1 else None))
    del phi
    npix = pts.get_new_pixel[0]
    n = pol * npix
    P = cm.SparseLO(npix, nt, pts._pix_dev, pol=pol, angle_processed=pts)
    F = cm.FilterLO(nt, [sub_len, sub_start], ns, args.ndet, pts._pix_dev)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A_local = P.T * F * N * F * P
    A = distributed.AllReduceLO(A_local) if world > 1 else A_local
    sky = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(99))
    d = P._apply(sky)
    d += 0.5 * torch.randn(nt, dtype=torch.float64, device="cuda", generator=g)
    b = P.T._apply(F._apply(N._apply(F._apply(d))))
    if world > 1:
        distributed.all_reduce_sum_(b)
    del d
    torch.cuda.synchronize()

    res = []
    t0 = time.perf_counter()
    x, info = cm.cg(A, b, M=Mbd, rtol=args.rtol, maxiter=args.maxiter, residuals=res)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    its = len(res) - 1 if info == 0 else len(res)
    rel = float(torch.linalg.norm(b - A._apply(x)) / torch.linalg.norm(b))
    # symmetry of the composed operator (F and N symmetric): <u, A v> = <v, A u>
    u = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    v = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(8))
    Av, Au = A._apply(v), A._apply(u)
    uav, vau = float(torch.dot(u, Av)), float(torch.dot(v, Au))
    sym_scale = float(torch.linalg.norm(u) * torch.linalg.norm(Av))

    # device time of the A apply alone (CUDA events; the TOD streams are far larger than L2)
    for _ in range(3):
        A._apply(x)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.time_iters):
        A._apply(x)
    e1.record()
    torch.cuda.synchronize()
    a_ms = torch.tensor([e0.elapsed_time(e1) / args.time_iters], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(a_ms, op=dist.ReduceOp.MAX)
    a_ms = float(a_ms.item())

    out = {"config": "configs[2]: Toeplitz noise (%d coefficients) + subscan offset filter, M_BD PCG" % args.nband,
           "world": world, "nt_total": nt * world, "nt_per_gpu": nt, "ndet_per_gpu": args.ndet, "npix": int(npix),
           "nside": args.nside, "nseg_per_gpu": F.nseg, "nband": args.nband,
           "plan": [type(f).__name__ for f in A_local.planned()],
           "cg": dict(info=int(info), iterations=its, seconds=dt, ms_per_iteration=1e3 * dt / max(its, 1),
                      true_relres=rel, rtol=args.rtol,
                      residual_first_last=[float(res[0]), float(res[-1])] if len(res) else None),
           "samples_per_s_per_pcg_iter": nt * world / (dt / max(its, 1)),
           "A_apply_ms": a_ms, "A_apply_samples_per_s": nt * world / (a_ms * 1e-3),
           "symmetry": {"u_Av": uav, "v_Au": vau, "rel": abs(uav - vau) / max(abs(uav), 1e-300),
                        "rel_to_norms": abs(uav - vau) / max(sym_scale, 1e-300),
                        "u_minus_v_norm": float(torch.linalg.norm(u - v))},
           "hbm_GB": torch.cuda.max_memory_allocated() / 1e9}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        if hasattr(A, "close"):
            A.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
