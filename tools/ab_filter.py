#!/usr/bin/env python
"""A/B of two builds of the library on ONE box: P^T F P apply time at configs[3]'s per-GPU share (development tool).
usage: python tools/ab_filter.py libA.so libB.so   -- runs A, B, A, B in fresh processes (CM2_LIB selects the build)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import workloads as wl
    nside, pol = 1024, 3
    nt, ns, pix, phi, sub_len, sub_start, g = wl.make_scan(int(5e8), nside, 1600, 800, 64, 8.0, seed=0)
    npix_full = 12 * nside ** 2
    pts = cm.ProcessTimeSamples(pix, npix_full, obspix=np.arange(npix_full), pol=pol, phi=phi)
    del phi
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, nt, pts._pix_dev, pol=pol, angle_processed=pts)
    F = cm.FilterLO(nt, [sub_len, sub_start], ns, 64, pts._pix_dev)
    A = P.T * F * P
    x = torch.randn(pol * npix, dtype=torch.float64, device="cuda")
    print("%s A_apply_ms=%.4f" % (os.path.basename(os.environ.get("CM2_LIB", "default")), wl.time_device(lambda: A._apply(x), 30)))


if __name__ == "__main__":
    if len(sys.argv) == 1:
        child()
    else:
        for lib in sys.argv[1:3] * 2:
            env = dict(os.environ, CM2_LIB=os.path.abspath(lib))
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
