#!/usr/bin/env python
"""Timing of the fused chains against the unfused chains of the same operators (development tool; CUDA
events, 1e8 samples generated on the device, inputs far larger than L2):
P^T T P for short Toeplitz bands (cm2_amatvec_toeplitz) and F P for the offset filter
(cm2_pointing_filter_mu)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, linearoperators as lo  # noqa: E402
from cosmomap2_b200.workloads import make_scan  # noqa: E402


def timeit(fn, reps=10, warm=3):
    if os.environ.get("KBENCH_REPS"):          # profiling runs: one launch per kernel
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
    lo.FUSE_TOEPLITZ_A = lo.FUSE_FILTER_P = True
    nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
    ndet, pol, nside = 64, 3, 512
    nt, ns, pix, phi, sub_len, sub_start, g = make_scan(nt, nside, 1000, 500, ndet, 8.0, seed=0)
    pts = cm.ProcessTimeSamples(pix, 12 * nside ** 2, obspix=np.arange(12 * nside ** 2), pol=pol, phi=phi)
    del phi
    npix = pts.get_new_pixel[0]
    n = pol * npix
    P = cm.SparseLO(npix, nt, pts._pix_dev, pol=pol, angle_processed=pts)
    F = cm.FilterLO(nt, [sub_len, sub_start], ns, ndet, pts._pix_dev)
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    out = {"nt": nt, "npix": int(npix), "pol": pol}
    alg = 20.0 * nt + 48.0 * npix
    for nband in (3, 9):
        N = cm.BlockLO(ns, synthetic.toeplitz_bands(ndet, nband), offdiag=True)
        A = P.T * N * P
        t = timeit(lambda: A._apply(x))
        assert isinstance(A.planned()[0], lo._FusedToeplitzA)
        y = A._apply(x)
        lo.fusion_enabled = False
        Ac = P.T * N * P
        tc = timeit(lambda: Ac._apply(x))
        yc = Ac._apply(x)
        lo.fusion_enabled = True
        out["toeplitz%d_fused_ms" % nband] = t
        out["toeplitz%d_fused_GBs" % nband] = alg / (t * 1e-3) / 1e9
        out["toeplitz%d_chain_ms" % nband] = tc
        out["toeplitz%d_fused_vs_chain_relerr" % nband] = float((y - yc).abs().max() / yc.abs().max())
        del N, A, Ac, y, yc
    FP = F * P
    t = timeit(lambda: FP._apply(x))
    assert FP.planned()[0]._runs
    d = FP._apply(x)
    tc = timeit(lambda: F._apply(P._apply(x)))
    dc = F._apply(P._apply(x))
    out["FP_fused_ms"] = t
    out["FP_fused_GBs"] = (28.0 * nt + 24.0 * npix) / (t * 1e-3) / 1e9
    out["FP_chain_ms"] = tc
    out["FP_fused_vs_chain_relerr"] = float((d - dc).abs().max() / dc.abs().max())
    if True:                                              # Legendre run-table path against the per-subscan kernel
        for order in (1, 3):
            Fk = cm.FilterLO(nt, [sub_len, sub_start], ns, ndet, pts._pix_dev, poly_order=order)
            res = {}
            for table in (False, True):
                lo.FILTER_POLY_RUN_TABLE = table
                Ak = P.T * Fk * P
                tk = timeit(lambda: Ak._apply(x))
                res[table] = Ak._apply(x)
                out["amatvec_leg%d_%s_ms" % (order, "table" if table else "subscan")] = tk
            lo.FILTER_POLY_RUN_TABLE = True
            out["amatvec_leg%d_table_vs_subscan_relerr" % order] = float((res[True] - res[False]).abs().max() / res[False].abs().max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
