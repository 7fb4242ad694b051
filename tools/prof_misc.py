#!/usr/bin/env python
"""A few launches of the fused Legendre A-matvec, the moments pass and the deflation kernels at a TOD
larger than L2 (ncu target; development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, _device as dv  # noqa: E402


def main():
    nt, pol = 40000000, 3
    sc = synthetic.raster_scan(nt, nside=512, ndet=64, nx=1000, ny=500, samples_per_pixel=8.0, seed=0,
                               with_data=False)
    nt = sc.nt
    sc.pix[np.random.default_rng(4).random(nt) < 0.01] = -1
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi)
    npix = pts.get_new_pixel[0]
    n = pol * npix
    P = cm.SparseLO(npix, nt, sc.pix, pol=pol, angle_processed=pts)
    x = dv.to_dev_f64(np.random.default_rng(1).standard_normal(n))
    F1 = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix, poly_order=1)
    F3 = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix, poly_order=3)
    A1, A3 = P.T * F1 * P, P.T * F3 * P
    r = 32
    Zt = torch.randn((r, n), dtype=torch.float64, device="cuda") / np.sqrt(n)
    Zd = cm.DeflationLO(Zt.t())
    for rep in range(3):
        A1._apply(x)
        A3._apply(x)
        pts._moments(npix)
        Zd.T._apply(x)
    torch.cuda.synchronize()
    print("ok", nt, npix)


if __name__ == "__main__":
    main()
