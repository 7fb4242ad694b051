#!/usr/bin/env python
"""Large-shard check (development tool): nt > 2^31 samples on one B200, pointing generated on the
device.  Verifies 64-bit indexing end to end (hit counts, P^T P 1 = counts, PCG in one iteration)
and reports the fused A-matvec bandwidth at that size."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import _device as dv  # noqa: E402


def main():
    nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2300000000
    nside, nx, ny, ndet, spp = 1024, 2000, 1000, 64, 8.0
    ns = nt // ndet
    nt = ns * ndet
    ring = 4 * nside
    dev = torch.device("cuda")
    t = torch.arange(ns, dtype=torch.int64, device=dev)
    sweep = int(nx * spp)
    isw = t // sweep
    frac = (t - isw * sweep).to(torch.float64) / sweep
    xpos = torch.where(isw % 2 == 0, frac, 1.0 - 1e-12 - frac) * nx
    pix = torch.empty(nt, dtype=torch.int32, device=dev)
    phi = torch.empty(nt, dtype=torch.float64, device=dev)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    w = 0.5 + torch.rand(ndet, generator=g, device=dev, dtype=torch.float64)
    for b in range(ndet):
        ix = torch.remainder(torch.floor(xpos + 0.01 * b * nx / ndet).to(torch.int64), nx)
        iy = torch.remainder(torch.floor(t.to(torch.float64) / ns * ny + 0.05 * b).to(torch.int64), ny)
        pix[b * ns:(b + 1) * ns] = ((2 * nside - ny // 2 + iy) * ring + (ring // 2 - nx // 2 + ix)).to(torch.int32)
        phi[b * ns:(b + 1) * ns] = 0.1 * b + 2 * np.pi * 2.5 / 200. * t.to(torch.float64) + \
            1e-3 * torch.randn(ns, generator=g, device=dev, dtype=torch.float64)
    del t, isw, frac, xpos
    npix_full = 12 * nside * nside
    N = cm.BlockLO(ns, w.cpu().numpy())
    pts = cm.ProcessTimeSamples(pix, npix_full, obspix=np.arange(npix_full), pol=3, phi=phi, w=N.diag)
    del phi
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, nt, pts._pix_dev, pol=3, angle_processed=pts)
    hits = P.hits()
    assert int(hits.sum()) == nt, (int(hits.sum()), nt)
    counts = np.asarray(pts.counts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=3)
    A = P.T * N * P
    x = dv.to_dev_f64(np.random.default_rng(0).standard_normal(3 * npix))
    y = A._apply(x)
    z = Mbd._apply(y)
    err = float((z - x).abs().max() / x.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        A._apply(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        A._apply(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    xs, info = cm.cg(A, y, M=Mbd, rtol=1e-10, maxiter=5)
    print(json.dumps({"nt": nt, "nt_gt_2^31": nt > 2 ** 31, "npix": int(npix), "hits_sum_ok": True,
                      "weighted_counts_sum": float(counts.sum()), "MbdA_minus_I_rel": err,
                      "amatvec_ms": ms, "amatvec_GBs": (20.0 * nt + 48.0 * npix) / (ms * 1e-3) / 1e9,
                      "cg_info": int(info), "cg_err": float((xs - x).abs().max() / x.abs().max()),
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
