#!/usr/bin/env python
"""How the fused A-matvec time depends on the clock state (development tool): the same kernel timed
after warm-ups of increasing length, with the SM clock / power sampled through NVML."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, _device as dv  # noqa: E402


def nvml():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    return lambda: (pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)


def main():
    pol = 3
    sc = synthetic.raster_scan(100000000, nside=512, ndet=64, nx=1000, ny=500, samples_per_pixel=8.0, seed=0,
                               with_data=False)
    nt = sc.nt
    N = cm.BlockLO(sc.ns, sc.weights)
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, nt, sc.pix, pol=pol, angle_processed=pts)
    x = dv.to_dev_f64(np.random.default_rng(1).standard_normal(pol * npix))
    d = dv.to_dev_f64(np.random.default_rng(2).standard_normal(nt))
    A = P.T * N * P
    A._apply(x)
    read = nvml()
    out = []
    for name, fn in (("amatvec_white", lambda: A._apply(x)), ("Pt", lambda: P.T._apply(d))):
        for warm in (3, 30, 300, 3000):
            torch.cuda.synchronize()
            import time
            time.sleep(1.0)                      # let the clocks fall back
            for _ in range(warm):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            clk = read()
            torch.cuda.synchronize()
            out.append(dict(kernel=name, warm=warm, ms=e0.elapsed_time(e1) / 20, sm_mhz=clk[0], power_w=clk[1]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
