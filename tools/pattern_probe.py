#!/usr/bin/env python
"""Pointing-pattern sensitivity of the fused white A-matvec (cm2_amatvec_white): samples per pixel crossing
1, 2, 4, 8, 32, a scan tilted 30 degrees against the pixel rows, and random pointing (the reference tests'
pairs_gen), each at 1e8 samples / IQU / nside 512, for the scatter variants (register run compression; staged
through shared memory with windows of 64 / 288 pixels).  CUDA events, inputs >> L2."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import workloads, _device as dv  # noqa: E402


def main():
    nt = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
    nside, nx, ny, ndet, pol = 512, 1000, 500, 64, 3
    out = []
    cases = [("spp=%g" % s, dict(spp=s)) for s in (1.0, 2.0, 4.0, 8.0, 32.0)] + [("tilt30 spp=8", dict(spp=8.0, tilt_deg=30.0)),
                                                                                ("random", None)]
    for name, kw in cases:
        if kw is None:
            pix, phi, g = workloads.random_pointing(nt, nside, nx, ny, seed=0)
            ntt, ns = nt, nt // ndet
        else:
            ntt, ns, pix, phi, _a, _b, g = workloads.make_scan(nt, nside, nx, ny, ndet, kw["spp"], seed=0, turnaround=0.0,
                                                               tilt_deg=kw.get("tilt_deg", 0.0))
        w = 0.5 + np.random.default_rng(0).random(ndet)
        N = cm.BlockLO(ns, w)
        pts = cm.ProcessTimeSamples(pix, 12 * nside ** 2, obspix=np.arange(12 * nside ** 2), pol=pol, phi=phi, w=N.diag)
        del phi
        npix = pts.get_new_pixel[0]
        P = cm.SparseLO(npix, ntt, pts._pix_dev, pol=pol, angle_processed=pts)
        A = P.T * N * P
        x = torch.randn(pol * npix, dtype=torch.float64, device="cuda", generator=g)
        row = {"pattern": name, "nt": ntt, "npix": int(npix), "algorithmic_bytes": 20.0 * ntt + 48.0 * npix}
        ref = None
        for label, wpix in (("registers", 0), ("staged64", 64), ("staged288", 288)):
            dv.call("cm2_amatvec_white_set_stage", wpix)
            y = A._apply(x).clone()
            if ref is None:
                ref = y
            else:
                row[label + "_vs_registers"] = float((y - ref).abs().max() / ref.abs().max())
            ms = workloads.time_device(lambda: A._apply(x), 20, warmup=5)
            row[label + "_ms"] = ms
            row[label + "_GBs"] = row["algorithmic_bytes"] / (ms * 1e-3) / 1e9
        dv.call("cm2_amatvec_white_set_stage", 0)
        # M_BD A = I check of the default variant (white noise, weights fed to M_BD)
        Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
        z = Mbd._apply(A._apply(x))
        row["MbdA_minus_I"] = float((z - x).abs().max() / x.abs().max())
        out.append(row)
        print(json.dumps(row), flush=True)
        del P, A, N, pts, pix, x, ref, y, z, Mbd
        torch.cuda.empty_cache()
    with open(os.path.join(ROOT, "gpurun_out", "pattern_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
