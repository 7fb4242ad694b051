#!/usr/bin/env python
"""
End-to-end two-level solve at scale (the shape of BASELINE.json configs[3]): offset-filtered
map-making  (P^T F P) x = P^T F d  on a synthetic raster scan, first with M_BD, then with the
two-level preconditioner M_2lvl built from a preconditioned-Arnoldi deflation space
(run_krypy_arnoldi -> find_ritz_eigenvalues -> CoarseLO(apply='eig') -> DeflationLO), exactly the
recipe of the reference's src/test_M2_precond_onto_real_data.py:54-122.

    python examples/solve_two_level.py --nt 5e8                       # one GPU
    torchrun --nproc-per-node 8 examples/solve_two_level.py --nt 5e8  # 4e9 samples on 8 GPUs

The pointing is generated on the device (inputs only).  With torchrun the TOD is sharded by
detector; the map-domain sums go through cosmomap2_b200.distributed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_scan(nt, nside, nx, ny, ndet, spp, seed, turnaround=0.05):
    dev = torch.device("cuda")
    ns = nt // ndet
    nt = ns * ndet
    ring = 4 * nside
    sweep = int(nx * spp / (1.0 - turnaround))
    t = torch.arange(ns, dtype=torch.int64, device=dev)
    isw = t // sweep
    frac = (t - isw * sweep).to(torch.float64) / sweep
    u = torch.clamp((frac - turnaround / 2) / (1.0 - turnaround), 0.0, 1.0 - 1e-12)
    xpos = torch.where(isw % 2 == 0, u, 1.0 - 1e-12 - u) * nx
    inside = (frac >= turnaround / 2) & (frac < 1.0 - turnaround / 2)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    pix = torch.empty(nt, dtype=torch.int32, device=dev)
    phi = torch.empty(nt, dtype=torch.float64, device=dev)
    for b in range(ndet):
        dx = (torch.rand(1, generator=g, device=dev).item() - 0.5) * 0.04 * nx
        dy = (torch.rand(1, generator=g, device=dev).item() - 0.5) * 0.1 * ny
        ix = torch.remainder(torch.floor(xpos + dx).to(torch.int64), nx)
        iy = torch.remainder(torch.floor(t.to(torch.float64) / ns * ny + dy).to(torch.int64), ny)
        p = ((2 * nside - ny // 2 + iy) * ring + (ring // 2 - nx // 2 + ix)).to(torch.int32)
        pix[b * ns:(b + 1) * ns] = torch.where(inside, p, torch.full_like(p, -1))
        phi[b * ns:(b + 1) * ns] = 3.0 * torch.rand(1, generator=g, device=dev).item() + \
            2 * np.pi * 2.5 / 200. * t.to(torch.float64) + 1e-3 * torch.randn(ns, generator=g, device=dev, dtype=torch.float64)
    nsweeps = int(ns // sweep)
    s0 = int(np.ceil(turnaround / 2 * sweep))
    s1 = int(np.ceil((1 - turnaround / 2) * sweep))
    sub_start = np.arange(nsweeps, dtype=np.int64) * sweep + s0
    sub_len = np.full(nsweeps, s1 - s0, dtype=np.int64)
    return nt, ns, pix, phi, sub_len, sub_start, g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=float, default=2e8, help="samples per GPU")
    ap.add_argument("--nside", type=int, default=1024)
    ap.add_argument("--nx", type=int, default=1600)
    ap.add_argument("--ny", type=int, default=800)
    ap.add_argument("--ndet", type=int, default=64)
    ap.add_argument("--r", type=int, default=32)
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
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import distributed, _device as dv

    pol = 3
    nt, ns, pix, phi, sub_len, sub_start, g = make_scan(int(args.nt), args.nside, args.nx, args.ny, args.ndet, 8.0, seed=rank)
    npix_full = 12 * args.nside ** 2
    pts = cm.ProcessTimeSamples(pix, npix_full, obspix=np.arange(npix_full), pol=pol, phi=phi,
                                comm=(True if world > 1 else None))
    del phi
    npix = pts.get_new_pixel[0]
    n = pol * npix
    P = cm.SparseLO(npix, nt, pts._pix_dev, pol=pol, angle_processed=pts)
    F = cm.FilterLO(nt, [sub_len, sub_start], ns, args.ndet, pts._pix_dev, poly_order=args.poly_order)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A_local = P.T * F * P
    A = distributed.AllReduceLO(A_local) if world > 1 else A_local
    # data: a random sky (same on every rank) seen through P, plus white noise and per-subscan offsets
    sky = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(99))
    d = P._apply(sky)
    d += 0.5 * torch.randn(nt, dtype=torch.float64, device="cuda", generator=g)
    b = P.T._apply(F._apply(d))
    if world > 1:
        distributed.all_reduce_sum_(b)
    del d
    torch.cuda.synchronize()

    def solve(M, label):
        res = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, info = cm.cg(A, b, M=M, rtol=args.rtol, maxiter=args.maxiter, residuals=res)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        its = len(res) - 1 if info == 0 else len(res)
        rel = float(torch.linalg.norm(b - A._apply(x)) / torch.linalg.norm(b))
        return x, dict(precond=label, info=int(info), iterations=its, seconds=dt, ms_per_iteration=1e3 * dt / max(its, 1),
                       true_relres=rel)

    out = {"world": world, "nt_total": nt * world, "nt_per_gpu": nt, "npix": int(npix), "nside": args.nside,
           "nseg_per_gpu": F.nseg, "poly_order": args.poly_order, "shard_m2": bool(args.shard_m2 and world > 1)}
    x_bd, out["M_BD"] = solve(Mbd, "M_BD")

    # ---- deflation space ---------------------------------------------------------------------------
    t0 = time.perf_counter()
    if args.coarse == "scan":
        # a-priori subdomain space from the scan order (cosmomap2_b200.scan_coarse_space): no Krylov phase
        Z = cm.scan_coarse_space(P, args.r, ns, A=A, Mbd=Mbd, smooth=args.smooth).t()
        r, m, theta, thr = args.r, 0, np.zeros(1), 0.0
    else:
        # the reference's recipe: preconditioned Arnoldi, Ritz vectors of the smallest Ritz values
        V, H, m = cm.run_krypy_arnoldi(A, torch.ones(n, dtype=torch.float64, device="cuda"), Mbd, 1e-5,
                                       maxiter=args.arnoldi, ortho="dmgs")
        theta = np.sort(np.linalg.eigvalsh(H[:H.shape[1], :]))
        r = min(args.r, len(theta) - 1)
        thr = 0.5 * (theta[r - 1] + theta[r])
        Z, r, th = cm.find_ritz_eigenvalues(H, V, threshold=thr, eigenvalues=True)
        del V
    Zc = Z.contiguous() if isinstance(Z, torch.Tensor) else Z
    AZ = torch.stack([A._apply(dv.to_dev_f64(Zc[:, i].contiguous())) for i in range(r)]).t()
    E = cm.CoarseLO(Zc, AZ, r, apply="eig")
    Zd, AZd = cm.DeflationLO(Zc), cm.DeflationLO(AZ)
    if args.shard_m2 and world > 1:
        M2 = distributed.ShardedTwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
    else:
        M2 = Mbd * (cm.lp.IdentityOperator(n) - AZd * E * Zd.T) + Zd * E * Zd.T  # fused at first use
    torch.cuda.synchronize()
    out["deflation"] = dict(kind=args.coarse, arnoldi_steps=int(m), r=int(r), ritz_min=float(theta[0]), ritz_cut=float(thr),
                            ritz_max=float(theta[-1]), discarded_E_modes=int(getattr(E, "ndiscarded", 0)),
                            build_seconds=time.perf_counter() - t0)
    x_m2, out["M_2lvl"] = solve(M2, "M_2lvl")
    # the two solutions agree where A sees them (P^T F P has the per-subscan-offset null space)
    ax1, ax2 = A._apply(x_bd), A._apply(x_m2)
    out["Ax_agreement"] = float(torch.linalg.norm(ax1 - ax2) / torch.linalg.norm(ax1))
    out["hbm_GB"] = torch.cuda.max_memory_allocated() / 1e9
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        if hasattr(A, "close"):
            A.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
