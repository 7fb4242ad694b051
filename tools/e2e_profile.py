import cProfile, pstats, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmomap2_b200 as cm
from cosmomap2_b200 import synthetic, _device as dv
sc = synthetic.config_c2(nt=int(os.environ.get("NT","20000000")), seed=0, with_data=False)
N = cm.BlockLO(sc.ns, sc.weights)
pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=3, phi=sc.phi, w=N.diag)
npix = pts.get_new_pixel[0]
P = cm.SparseLO(npix, sc.nt, sc.pix, pol=3, angle_processed=pts)
Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=3)
A = P.T * N * P
n = 3 * npix
b = dv.pinned_array(n); b[...] = np.random.default_rng(0).standard_normal(n)
for _ in range(3):
    cm.cg(A, b, M=Mbd, rtol=1e-30, maxiter=1)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    cm.cg(A, b, M=Mbd, rtol=1e-30, maxiter=1)
torch.cuda.synchronize()
print("per call ms", (time.perf_counter() - t) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    cm.cg(A, b, M=Mbd, rtol=1e-30, maxiter=1)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
