#!/usr/bin/env python
"""Two launches each of the round-2 hot kernels at TODs larger than L2 (ncu target; development tool):
fused white A-matvec (register / interleaved / staged / pixel-sorted scatter), offset-filtered A-matvec,
FFT Toeplitz with 4096 coefficients, two-level preconditioner apply, cooperative PCG tail."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import workloads, synthetic, linearoperators as lo  # noqa: E402


def white(nt, nside, nx, ny, spp, reps=2, **kw):
    ndet, pol = 64, 3
    ntt, ns, pix, phi, sl, ss, g = workloads.make_scan(nt, nside, nx, ny, ndet, spp, seed=0, **kw)
    N = cm.BlockLO(ns, 0.5 + np.random.default_rng(0).random(ndet))
    pts = cm.ProcessTimeSamples(pix, 12 * nside ** 2, obspix=np.arange(12 * nside ** 2), pol=pol, phi=phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, ntt, pts._pix_dev, pol=pol, angle_processed=pts)
    x = torch.randn(pol * npix, dtype=torch.float64, device="cuda", generator=g)
    return ntt, ns, ndet, pts, P, N, x, sl, ss


def main():
    # configs[1] shape, 4e7 samples: register scatter
    ntt, ns, ndet, pts, P, N, x, sl, ss = white(40000000, 512, 1000, 500, 8.0, turnaround=0.0)
    A = P.T * N * P
    for _ in range(2):
        A._apply(x)
    # the PCG tail on the same problem
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, P.ncols, pol=3)
    b = A._apply(x)
    cm.cg(A, b, M=Mbd, rtol=1e-30, maxiter=2)
    # 4 samples per pixel: staged scatter; 1 sample per pixel: pixel-sorted pointing
    for spp in (4.0, 1.0):
        _ntt, _ns, _nd, _pts, P2, N2, x2, _a, _b = white(40000000, 512, 1000, 500, spp, turnaround=0.0)
        A2 = P2.T * N2 * P2
        for _ in range(2):
            A2._apply(x2)
        del P2, N2, A2, x2, _pts
    # nside 2048 patch: detector-interleaved order
    ntt4, ns4, _nd, pts4, P4, N4, x4, _a, _b = white(200000000, 2048, 3200, 1600, 8.0, turnaround=0.0)
    A4 = P4.T * N4 * P4
    for _ in range(2):
        A4._apply(x4)
    del P4, N4, A4, x4, pts4
    torch.cuda.empty_cache()
    # offset-filtered A-matvec (configs[3] kernel) on an nside 1024 patch
    ntt3, ns3, nd3, pts3, P3, _N3, x3, sl3, ss3 = white(100000000, 1024, 1600, 800, 8.0)
    F3 = cm.FilterLO(ntt3, [sl3, ss3], ns3, nd3, pts3._pix_dev)
    A3 = P3.T * F3 * P3
    for _ in range(3):
        A3._apply(x3)
    # two-level apply, r = 32
    Mbd3 = cm.BlockDiagonalPreconditionerLO(pts3, P3.ncols, pol=3)
    Zt = cm.scan_coarse_space(P3, 32, ns3)
    AZt = torch.stack([A3._apply(Zt[i]) for i in range(4)] + [Zt[i] for i in range(4, 32)])
    E = cm.CoarseLO(Zt.t(), AZt.t(), 32, apply="eig")
    Zd, AZd = cm.DeflationLO(Zt.t()), cm.DeflationLO(AZt.t())
    M2 = Mbd3 * (cm.lp.IdentityOperator(3 * P3.ncols) - AZd * E * Zd.T) + Zd * E * Zd.T
    for _ in range(2):
        M2._apply(x3)
    del P3, F3, A3, x3, pts3, Zd, AZd, Zt, AZt, M2
    torch.cuda.empty_cache()
    # FFT Toeplitz, 4096 coefficients, 8 detectors x 2.5e6 samples
    nt2, nd2 = 20000000, 8
    N2 = cm.BlockLO(nt2 // nd2, synthetic.toeplitz_bands(nd2, 4096), offdiag=True)
    d = torch.randn(nt2, dtype=torch.float64, device="cuda")
    for _ in range(2):
        N2._apply(d)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
