#!/usr/bin/env python
"""A few launches of the offset-filtered A-matvec P^T F P (k_seg_mean over the run table + k_amatvec_filter_mu, one
pass over the TOD) on a raster scan larger than L2 -- configs[3]'s kernels (ncu target; development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, _device as dv  # noqa: E402


def main():
    nt, pol = 100000000, 3
    sc = synthetic.raster_scan(nt, nside=512, ndet=64, nx=1000, ny=500, samples_per_pixel=8.0, seed=0,
                               with_data=False)
    nt = sc.nt
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, nt, sc.pix, pol=pol, angle_processed=pts)
    F = cm.FilterLO(nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix)
    A = P.T * F * P
    x = dv.to_dev_f64(np.random.default_rng(1).standard_normal(pol * npix))
    for rep in range(4):
        A._apply(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(10):
        A._apply(x)
    e1.record()
    torch.cuda.synchronize()
    print("ok nt=%d npix=%d nseg=%d A_apply_ms=%.4f" % (nt, npix, len(sc.sub_len) * sc.ndet, e0.elapsed_time(e1) / 10))


if __name__ == "__main__":
    main()
