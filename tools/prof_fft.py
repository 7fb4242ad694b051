#!/usr/bin/env python
"""Two launches of the FFT Toeplitz kernel at 4096 coefficients with 16384-sample windows and with 32768-sample
windows on 2-CTA clusters, 8 detectors x 2.5e6 samples (ncu target; development tool)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cosmomap2_b200 as cm  # noqa: E402
from cosmomap2_b200 import synthetic, linearoperators as lo  # noqa: E402

nt, nd = 20000000, 8
d = torch.randn(nt, dtype=torch.float64, device="cuda")
for L, pair_min in ((4096, 10 ** 9), (4096, 2000)):
    lo.TOEPLITZ_FFT_PAIR_MIN_BAND = pair_min
    N = cm.BlockLO(nt // nd, synthetic.toeplitz_bands(nd, L), offdiag=True)
    for _ in range(2):
        N._apply(d)
torch.cuda.synchronize()
print("ok")
