import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmomap2_b200 as cm
from cosmomap2_b200 import synthetic, _device as dv, linearoperators as lo
from kbench import timeit
sc = synthetic.raster_scan(100000000, nside=512, ndet=64, nx=1000, ny=500, samples_per_pixel=8.0, seed=0, with_data=False)
pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=3, phi=sc.phi)
npix = pts.get_new_pixel[0]
P = cm.SparseLO(npix, sc.nt, sc.pix, pol=3, angle_processed=pts)
F = cm.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix)
x = dv.to_dev_f64(np.random.default_rng(1).standard_normal(3 * npix))
A = P.T * F * P
A._apply(x)
fa = [f for f in A.planned() if isinstance(f, lo._FusedFilterA)][0]
rt = fa._runs
y = dv.empty_f64(3 * npix)
st = dv.stream
t1 = timeit(lambda: dv.call("cm2_filter_seg_mean", dv.ptr(rt["run_pix"]), dv.ptr(rt["run_mom"]), dv.ptr(rt["seg_first"]),
                            dv.ptr(rt["seg_nruns"]), F.nseg, 3, dv.ptr(x), dv.ptr(rt["mu"]), st()))
t2 = timeit(lambda: dv.call("cm2_amatvec_filter_mu", dv.ptr(P._pix_dev), dv.ptr(P._cos_dev), dv.ptr(P._sin_dev), P.nrows, 3,
                            dv.ptr(F._seg_start), dv.ptr(F._seg_end), dv.ptr(rt["mu"]), dv.ptr(rt["tile_seg"]), dv.ptr(rt["tile_flag"]), F.nseg,
                            dv.ptr(x), dv.ptr(y), P.ncols, 1, st()))
print(json.dumps({"nseg": F.nseg, "nruns": rt["nruns"], "seg_mean_ms": t1, "amatvec_filter_mu_ms": t2}))
