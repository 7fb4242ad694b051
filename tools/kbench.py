#!/usr/bin/env python
"""Per-kernel timing at the bench workload (development tool; CUDA events, L2-exceeding inputs)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, _device as dv  # noqa: E402


def timeit(fn, reps=20, warm=3):
    if os.environ.get("KBENCH_REPS"):          # profiling runs: one or two launches per kernel
        reps, warm = int(os.environ["KBENCH_REPS"]), 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=100000000)
    ap.add_argument("--spp", type=float, default=8.0)
    ap.add_argument("--random", action="store_true", help="random pointing instead of a raster scan")
    ap.add_argument("--pol", type=int, default=3)
    args = ap.parse_args()
    pol = args.pol
    sc = synthetic.raster_scan(args.nt, nside=512, ndet=64, nx=1000, ny=500, samples_per_pixel=args.spp, seed=0,
                               with_data=False)
    if args.random:
        sc.pix = np.random.default_rng(0).permutation(sc.pix)
    nt = sc.nt
    N = cm.BlockLO(sc.ns, sc.weights)
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, nt, sc.pix, pol=pol, angle_processed=pts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    F = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix)
    n = pol * npix
    x = dv.to_dev_f64(np.random.default_rng(1).standard_normal(n))
    d = dv.to_dev_f64(np.random.default_rng(2).standard_normal(nt))
    A = P.T * N * P
    AF = P.T * F * P
    out = {"nt": nt, "npix": int(npix), "pol": pol}
    for _ in range(1500):                       # ~0.5 s of load: clocks settled before the first timing
        A._apply(x)
    torch.cuda.synchronize()
    gb = lambda b, ms: b / (ms * 1e-3) / 1e9  # noqa: E731
    bpp = 4 + (16 if pol > 1 else 0)
    t = timeit(lambda: A._apply(x)); out["amatvec_white_ms"] = t; out["amatvec_white_GBs"] = gb(bpp * nt + 16 * n, t)
    t = timeit(lambda: AF._apply(x)); out["amatvec_filter_ms"] = t; out["amatvec_filter_GBs"] = gb(bpp * nt + 16 * n, t)
    t = timeit(lambda: P._apply(x)); out["P_ms"] = t; out["P_GBs"] = gb((bpp + 8) * nt + 8 * n, t)
    t = timeit(lambda: P.T._apply(d)); out["Pt_ms"] = t; out["Pt_GBs"] = gb((bpp + 8) * nt + 8 * n, t)
    t = timeit(lambda: N._apply(d)); out["Nwhite_ms"] = t; out["Nwhite_GBs"] = gb(16 * nt, t)
    t = timeit(lambda: F._apply(d)); out["F_ms"] = t; out["F_GBs"] = gb(20 * nt, t)
    # Legendre subscan filter (standalone and fused into the A-matvec), ground filter
    from cosmomap2_b200 import linearoperators as lo
    lo.FILTER_STAGED = False
    t = timeit(lambda: F._apply(d)); out["F_first_kernel_ms"] = t; out["F_first_kernel_GBs"] = gb(20 * nt, t)
    lo.FILTER_STAGED = True
    pixf = np.array(sc.pix, copy=True)
    pixf[np.random.default_rng(4).random(nt) < 0.01] = -1
    from cosmomap2_b200 import _cabi
    _cabi.call("cm2_filter_poly_set_tma", 0)
    t = timeit(lambda: F._apply(d)); out["F_ldg_ms"] = t; out["F_ldg_GBs"] = gb(20 * nt, t)
    _cabi.call("cm2_filter_poly_set_tma", 1)
    t = timeit(lambda: F._apply(d)); out["F_tma_ms"] = t; out["F_tma_GBs"] = gb(20 * nt, t)
    for order in (1, 3):
        Fk = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix, poly_order=order)
        _cabi.call("cm2_filter_poly_set_tma", 0)
        t = timeit(lambda: Fk._apply(d)); out["F_leg%d_ldg_ms" % order] = t
        _cabi.call("cm2_filter_poly_set_tma", 1)
        t = timeit(lambda: Fk._apply(d)); out["F_leg%d_ms" % order] = t; out["F_leg%d_GBs" % order] = gb(20 * nt, t)
        Ff = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pixf, poly_order=order)
        t = timeit(lambda: Ff._apply(d)); out["F_leg%d_flagged_ms" % order] = t
        out["F_leg%d_flagged_GBs" % order] = gb(20 * nt, t)
        AFk = P.T * Fk * P
        t = timeit(lambda: AFk._apply(x)); out["amatvec_leg%d_ms" % order] = t
        out["amatvec_leg%d_GBs" % order] = gb(bpp * nt + 16 * n, t)
    del pixf, Ff
    ground = ((np.arange(nt, dtype=np.int64) % sc.ns) // 50) % 400
    ground[np.random.default_rng(5).random(nt) < 0.01] = -1
    Gf = cm.GroundFilterLO(ground)
    del ground
    t = timeit(lambda: Gf._apply(d)); out["ground_ms"] = t; out["ground_GBs"] = gb(32 * nt, t)
    del Gf
    t = timeit(lambda: Mbd._apply(x)); out["Mbd_ms"] = t; out["Mbd_GBs"] = gb(48 * npix + 16 * n, t)
    t = timeit(lambda: pts._moments(npix)); out["moments_ms"] = t
    bands = synthetic.toeplitz_bands(64, 64)
    NT = cm.BlockLO(sc.ns, bands, offdiag=True)
    t = timeit(lambda: NT._apply(d), reps=5, warm=1); out["toeplitz64_ms"] = t
    out["toeplitz64_GFLOPs"] = 2.0 * (2 * 64 - 1) * nt / (t * 1e-3) / 1e9
    for L in (3, 9):                            # the reference tests' composition: P^T T P in one TOD pass
        ATs = P.T * cm.BlockLO(sc.ns, synthetic.toeplitz_bands(64, L), offdiag=True) * P
        t = timeit(lambda: ATs._apply(x)); out["amatvec_toeplitz%d_ms" % L] = t
        out["amatvec_toeplitz%d_GBs" % L] = gb(bpp * nt + 16 * n, t)
    FP = F * P                                  # offset filter fused into the gather
    t = timeit(lambda: FP._apply(x)); out["FP_fused_ms"] = t; out["FP_fused_GBs"] = gb((bpp + 8) * nt + 8 * n, t)
    for L in (256, 4096):
        NTf = cm.BlockLO(sc.ns, synthetic.toeplitz_bands(64, L), offdiag=True)
        t = timeit(lambda: NTf._apply(d), reps=3, warm=1)
        out["toeplitz%d_fft_ms" % L] = t
        out["toeplitz%d_fft_Msamples_s" % L] = nt / (t * 1e-3) / 1e6
        out["toeplitz%d_equiv_direct_GFLOPs" % L] = 2.0 * (2 * L - 1) * nt / (t * 1e-3) / 1e9
    # deflation / two-level preconditioner at r = 32 (rows a10-a13): 8*r*n bytes per pass over Z
    r = 32
    Zt = torch.randn((r, n), dtype=torch.float64, device="cuda") / np.sqrt(n)
    AZt = torch.stack([A._apply(Zt[i]) for i in range(r)])
    Zd, AZd = cm.DeflationLO(Zt.t()), cm.DeflationLO(AZt.t())
    E = cm.CoarseLO(Zt.t(), AZt.t(), r, apply="eig")
    M2 = cm.TwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
    yv = dv.to_dev_f64(np.random.default_rng(3).standard_normal(r))
    t = timeit(lambda: Zd.T._apply(x)); out["Zt_x_ms"] = t; out["Zt_x_GBs"] = gb(8.0 * r * n, t)
    t = timeit(lambda: Zd._apply(yv)); out["Z_y_ms"] = t; out["Z_y_GBs"] = gb(8.0 * r * n, t)
    t = timeit(lambda: M2._apply(x)); out["M2_ms"] = t; out["M2_GBs"] = gb(3 * 8.0 * r * n + 48.0 * npix + 16.0 * n, t)
    t = timeit(lambda: cm.CoarseLO(Zt.t(), AZt.t(), r, apply="eig"), reps=3, warm=1); out["coarse_build_ms"] = t
    from cosmomap2_b200.pcg import PCG
    b = P.T._apply(N._apply(d))
    solver = PCG(A, Mbd, n)

    def step():
        solver.start(b)
        solver.step_async()
        solver.tick()
    t = timeit(step, reps=50, warm=5); out["pcg_iter_ms"] = t
    print(json.dumps(out))


if __name__ == "__main__":
    main()
