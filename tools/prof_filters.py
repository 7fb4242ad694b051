#!/usr/bin/env python
"""A few launches of the SURVEY 8(f) kernels at a TOD larger than L2 (ncu target; development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, _device as dv  # noqa: E402


def main():
    nt = int(os.environ.get("PROF_NT", 40000000))
    pol = 3
    sc = synthetic.raster_scan(nt, nside=512, ndet=64, nx=1000, ny=500, samples_per_pixel=8.0, seed=0,
                               with_data=False)
    nt = sc.nt
    sc.pix[np.random.default_rng(4).random(nt) < 0.01] = -1
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, nt, sc.pix, pol=pol, angle_processed=pts)
    x = dv.to_dev_f64(np.random.default_rng(1).standard_normal(pol * npix))
    d = dv.to_dev_f64(np.random.default_rng(2).standard_normal(nt))
    ground = ((np.arange(nt, dtype=np.int64) % sc.ns) // 50) % 400
    Gf = cm.GroundFilterLO(ground)
    ops = []
    for order in (0, 1, 3):
        F = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix, poly_order=order)
        ops.append((F, d))
        if order:
            ops.append((P.T * F * P, x))
    ops.append((Gf, d))
    for rep in range(3):
        for op, v in ops:
            op._apply(v)
    torch.cuda.synchronize()
    print("ok", nt, npix)


if __name__ == "__main__":
    main()
