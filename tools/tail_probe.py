#!/usr/bin/env python
"""Device time of the PCG tail kernels (cm2_pcg_bd_iter, cm2_pcg_bd_reset, generic-M updates) at the pixel counts of
configs[1], [3], [4] (development tool)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402,F401
from cosmomap2_b200 import workloads, _device as dv  # noqa: E402

for npix in (500000, 1280000, 5120000):
    n = 3 * npix
    g = torch.Generator(device="cuda").manual_seed(1)
    inv = torch.rand(6 * npix, dtype=torch.float64, device="cuda", generator=g) + 0.5
    p, q, x, r, z, b = (torch.randn(n, dtype=torch.float64, device="cuda", generator=g) for _ in range(6))
    scal = dv.zeros_f64(16)
    st = dv.stream

    def reset():
        dv.call("cm2_pcg_bd_reset", dv.ptr(inv), npix, 3, dv.ptr(r), dv.ptr(z), dv.ptr(scal), 0.0, 0.0, dv.ptr(b), dv.ptr(x),
                dv.ptr(p), st())

    def it():
        scal[7] = 0.0
        dv.call("cm2_pcg_bd_iter", dv.ptr(inv), npix, 3, dv.ptr(p), dv.ptr(q), dv.ptr(x), dv.ptr(r), dv.ptr(z), dv.ptr(scal), st())

    def upd():
        dv.call("cm2_pcg_update_p", dv.ptr(r), dv.ptr(z), dv.ptr(p), n, dv.ptr(scal), st())
        dv.call("cm2_pcg_update_xr", dv.ptr(p), dv.ptr(q), dv.ptr(x), dv.ptr(r), n, dv.ptr(scal), st())

    reset()
    out = {"npix": npix, "bd_reset_ms": workloads.time_device(reset, 50), "bd_iter_ms": workloads.time_device(it, 50),
           "generic_update_p_plus_xr_ms": workloads.time_device(upd, 50),
           "bd_iter_bytes": 336.0 * npix, "bd_reset_bytes": 168.0 * npix}
    out["bd_iter_GBs"] = out["bd_iter_bytes"] / out["bd_iter_ms"] / 1e6
    print(json.dumps(out), flush=True)
