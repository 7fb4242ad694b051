"""Test infrastructure for the multi-GPU path (not collected by pytest: no ``test_`` prefix).

* ``EmulatedShardedPCG``: the pixel-sharded PCG kernel (cm2_pcg_bd_sharded) with ALL ranks played by one
  cooperative launch on one GPU (``nvirt = world``): the same device code, flags included, without
  NVLink -- what a box with a single GPU can check.
* ``python -m torch.distributed.run --nproc-per-node N tests/multirank.py <case> <out.json>``: worker of
  the N-rank tests in tests/test_gpu_multirank.py (one process per GPU, NCCL).
"""
import ctypes
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class EmulatedShardedPCG(object):
    def __init__(self, A_locals, Mbd, timeout_s=5.0):
        import torch
        from cosmomap2_b200 import _device as dv, distributed
        self.dv, self.torch = dv, torch
        self.A_locals, self.world = list(A_locals), len(A_locals)
        self.pol, self.npix = Mbd.pol, Mbd._n
        self.n = self.pol * self.npix
        self.pix_lo = distributed.partition_pixels(self.npix, self.world)
        W = self.world
        self.y = [dv.zeros_f64(self.n) for _ in range(W)]
        self.p = [dv.zeros_f64(self.n) for _ in range(W)]
        nsig = int(dv.call("cm2_pcg_sharded_signal_bytes"))
        self.sig = [torch.zeros(nsig, dtype=torch.uint8, device=dv.device()) for _ in range(W)]
        self.sl = [(self.pol * self.pix_lo[g], self.pol * self.pix_lo[g + 1]) for g in range(W)]
        mk = lambda: [dv.zeros_f64(max(b - a, 2)) for a, b in self.sl]
        self.x, self.r, self.z, self.q = mk(), mk(), mk(), mk()
        self.inv = [Mbd._inv_dev[6 * self.pix_lo[g]:6 * self.pix_lo[g + 1]].clone() if self.pix_lo[g + 1] > self.pix_lo[g]
                    else dv.zeros_f64(6) for g in range(W)]
        self.scal = [dv.zeros_f64(16) for _ in range(W)]
        self.part = [dv.empty_f64(int(dv.call("cm2_pcg_sharded_work_doubles"))) for _ in range(W)]
        self.gen = 0
        self.timeout_s = timeout_s
        arr = ctypes.c_void_p * W
        self._arr = arr
        self._tab = dict((k, arr(*[t.data_ptr() for t in getattr(self, k)]))
                         for k in ("y", "p", "sig", "x", "r", "z", "q", "inv", "scal", "part"))
        self._pix_lo = (ctypes.c_int64 * (W + 1))(*self.pix_lo)

    def _launch(self, reset, b_tab, atol=0.0, rtol=0.0):
        t = self._tab
        self.gen += 1
        self.dv.call("cm2_pcg_bd_sharded", 1 if reset else 0, self.pol, self.world, 0, self.world, self._pix_lo,
                     t["y"], t["p"], t["sig"], t["x"], t["r"], t["z"], t["q"], t["inv"], b_tab, t["scal"], t["part"],
                     self.gen, float(atol), float(rtol), self.timeout_s, self.dv.stream())

    def start(self, b, atol=0.0, rtol=0.0):
        self._b = [b[a:bb] if bb > a else self.z[g] for g, (a, bb) in enumerate(self.sl)]
        self._launch(True, self._arr(*[t.data_ptr() for t in self._b]), atol, rtol)

    def step(self):
        for g, A in enumerate(self.A_locals):
            with self.dv.map_output(self.y[g]):
                yy = A._apply(self.p[g])
            self.dv.land(yy, self.y[g])
        self._launch(False, None)

    def solution(self):
        return self.torch.cat([self.x[g][:b - a] for g, (a, b) in enumerate(self.sl)])

    def scalars(self):
        return [s.cpu().numpy().copy() for s in self.scal]


def split_problem(sc, pol, world, cm, correlated=False):
    """One global set-up (pixel set, M_BD) + ``world`` local operators A_g = P_g^T N_g P_g over contiguous
    detector shards of the scan ``sc``; returns (A_global, A_locals, Mbd, b, npix).  ``correlated``: N is
    a 3-coefficient Toeplitz band per detector (a solve with real CG dynamics; white noise converges in
    one iteration)."""
    from cosmomap2_b200 import distributed, synthetic
    pix = sc.pix.astype(np.int64)
    Nw = cm.BlockLO(sc.ns, sc.weights)
    bands = synthetic.toeplitz_bands(sc.ndet, 3, seed=9) if correlated else None
    N = cm.BlockLO(sc.ns, bands, offdiag=True) if correlated else Nw
    pts = cm.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=Nw.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A = P.T * N * P
    b = P.T * (N * sc.d)
    A_locals = []
    for g in range(world):
        lo, hi = distributed.shard_detectors(sc.ndet, world, g)
        a, e = lo * sc.ns, hi * sc.ns
        ang = types.SimpleNamespace(cos=np.asarray(pts.cos)[a:e].copy(), sin=np.asarray(pts.sin)[a:e].copy()) if pol > 1 else None
        Pg = cm.SparseLO(npix, e - a, pix[a:e].copy(), pol=pol, angle_processed=ang)
        Ng = cm.BlockLO(sc.ns, bands[lo:hi], offdiag=True) if correlated else cm.BlockLO(sc.ns, sc.weights[lo:hi])
        A_locals.append(Pg.T * Ng * Pg)
    return A, A_locals, Mbd, b, npix


# =================================================================================================
# torchrun worker
# =================================================================================================
def _rank_problem(cm, distributed, synthetic, rank, world, pol=3, nt=400000, ndet=8, correlated=False):
    """Every rank generates the same scan and keeps its detectors; the pixel set comes from the summed
    moments (ProcessTimeSamples(comm=True))."""
    sc = synthetic.raster_scan(nt, nside=64, ndet=ndet, nx=90, ny=50, samples_per_pixel=6.0, seed=5,
                               flag_turnarounds=True)
    (pix, phi, d), (lo, hi) = distributed.shard_tod([sc.pix.astype(np.int64), sc.phi, sc.d], sc.ndet, sc.ns, world, rank)
    pix = pix.copy()
    w = sc.weights[lo:hi]
    N = cm.BlockLO(sc.ns, w)
    pts = cm.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=phi, w=N.diag, comm=True)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, len(pix), pix, pol=pol, angle_processed=pts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    if correlated:
        bands = synthetic.toeplitz_bands(hi - lo, 3, seed=3 + lo)
        Nn = cm.BlockLO(sc.ns, bands, offdiag=True)
    else:
        Nn = N
    A_loc = P.T * Nn * P
    b = P.T._apply(Nn._apply(cm._device.to_dev_f64(d)))
    distributed.all_reduce_sum_(b)
    return sc, A_loc, Mbd, b, npix


def _single_gpu_reference(cm, synthetic, pol=3, nt=400000, ndet=8):
    """The same problem solved by ONE process holding all detectors (the thing the sharded run must equal)."""
    sc = synthetic.raster_scan(nt, nside=64, ndet=ndet, nx=90, ny=50, samples_per_pixel=6.0, seed=5,
                               flag_turnarounds=True)
    pix = sc.pix.astype(np.int64)
    N = cm.BlockLO(sc.ns, sc.weights)
    pts = cm.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A = P.T * N * P
    b = P.T * (N * sc.d)
    return sc, A, Mbd, b, npix


def worker(case, out_path):
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import distributed, synthetic
    from cosmomap2_b200 import _device as dv
    res = {"case": case, "world": world}

    def same_on_ranks(t):
        ts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(ts, t.contiguous())
        return all(torch.equal(ts[0], u) for u in ts)

    if case == "p2p_allreduce":
        ok = True
        for n in (150001, 150000, 7, 1):
            ar = distributed.P2PAllReduce(n)
            g = torch.Generator(device="cuda")
            g.manual_seed(100 + rank)
            y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
            allv = [torch.empty_like(y) for _ in range(world)]
            dist.all_gather(allv, y)
            ref = allv[0].clone()
            for k in range(1, world):
                ref += allv[k]
            nccl = y.clone()
            dist.all_reduce(nccl)
            for it in range(3):
                out = ar(y).clone()
                ok &= bool(torch.equal(out, ref)) and same_on_ranks(out)
                ok &= bool(((out - nccl).abs().max() <= 1e-12 * nccl.abs().max()).item())
            ok &= ar.error() == 0
            ar.close()
        res["ok"] = ok

    elif case in ("sharded_solve", "sharded_solve_pol1"):
        pol = 1 if case.endswith("pol1") else 3
        sc, A_loc, Mbd, b, npix = _rank_problem(cm, distributed, synthetic, rank, world, pol=pol)
        A = distributed.AllReduceLO(A_loc)
        solver = A.sharded_solver(Mbd)
        res["sharded_available"] = solver is not None
        r_sh, r_rep = [], []
        x_sh, info_sh = cm.cg(A, b, M=Mbd, rtol=1e-12, maxiter=60, residuals=r_sh)
        res["used_sharded"] = A._p2p is not None and len(A._sharded) == 1
        # replicated path on the same operators (all-reduce + cooperative tail on every rank)
        os.environ["CM2_SHARDED_PCG"] = "0"
        x_rep, info_rep = cm.cg(A, b, M=Mbd, rtol=1e-12, maxiter=60, residuals=r_rep)
        os.environ["CM2_SHARDED_PCG"] = "1"
        # slices: gather="shard" returns this rank's part of the same solution
        x_slice, _ = cm.cg(A, b, M=Mbd, rtol=1e-12, maxiter=60, gather="shard")
        lo = distributed.partition_pixels(npix, world)
        # a second solve: the fp64 reductions of the TOD pass land in a run-dependent order, so the two solutions
        # agree to rounding (1e-10 of the solution, as every other comparison here), not always bit for bit
        mine = x_sh[pol * lo[rank]:pol * lo[rank + 1]]
        res["slice_shape_ok"] = tuple(x_slice.shape) == tuple(mine.shape)
        res["slice_rel"] = float((x_slice - mine).abs().max() / x_sh.abs().max()) if mine.numel() else 0.0
        res["slice_equal"] = res["slice_shape_ok"] and res["slice_rel"] < 1e-10
        res["slice_bit_equal"] = bool(torch.equal(x_slice, mine))
        res["identical_on_ranks"] = same_on_ranks(x_sh)
        res["info"] = [int(info_sh), int(info_rep)]
        res["iters"] = [len(r_sh), len(r_rep)]
        res["x_rel"] = float((x_sh - x_rep).abs().max() / x_rep.abs().max())
        bn = float(torch.linalg.norm(b))
        m = min(len(r_sh), len(r_rep))
        res["res_rel"] = float(np.max(np.abs(np.array(r_sh[:m]) - np.array(r_rep[:m]))) / bn)
        # ... and the one-process solve of the whole scan + the oracle
        sc1, A1, M1, b1, npix1 = _single_gpu_reference(cm, synthetic, pol=pol)
        r1 = []
        x1, info1 = cm.cg(A1, b1, M=M1, rtol=1e-12, maxiter=60, residuals=r1)
        res["npix_equal"] = int(npix1) == int(npix)
        res["x_rel_single"] = float(np.max(np.abs(dv.to_host(x_sh) - x1)) / np.max(np.abs(x1)))
        res["iters_single"] = len(r1)
        if rank == 0:
            import scipy.sparse.linalg as spla
            import oracle
            pix = sc1.pix.astype(np.int64)
            No = oracle.BlockLO(sc1.ns, sc1.weights)
            pts = oracle.ProcessTimeSamples(pix, sc1.npix_full, pol=pol, phi=sc1.phi, w=No.diag)
            Po = oracle.SparseLO(pts.get_new_pixel[0], sc1.nt, pix, pol=pol, angle_processed=pts)
            Mo = oracle.BlockDiagonalPreconditionerLO(pts, pts.get_new_pixel[0], pol=pol)
            Ao = Po.T * No * Po
            bo = Po.T * (No * sc1.d)
            it = [0]
            xo, info_o = spla.cg(Ao, bo, M=Mo, rtol=1e-12, maxiter=60, callback=lambda xk: it.__setitem__(0, it[0] + 1))
            res["x_rel_oracle"] = float(np.max(np.abs(dv.to_host(x_sh) - xo)) / np.max(np.abs(xo)))
            res["iters_oracle"] = it[0] + 1
        A.check()
        A.close()

    elif case == "sharded_solve_toeplitz":
        # a solve that takes many iterations: short-band Toeplitz noise, fused P^T T P per rank
        sc, A_loc, Mbd, b, npix = _rank_problem(cm, distributed, synthetic, rank, world, pol=3, correlated=True)
        A = distributed.AllReduceLO(A_loc)
        r_sh, r_rep = [], []
        x_sh, info_sh = cm.cg(A, b, M=Mbd, rtol=1e-10, maxiter=200, residuals=r_sh)
        res["used_sharded"] = len(A._sharded) == 1
        os.environ["CM2_SHARDED_PCG"] = "0"
        x_rep, info_rep = cm.cg(A, b, M=Mbd, rtol=1e-10, maxiter=200, residuals=r_rep)
        os.environ["CM2_SHARDED_PCG"] = "1"
        res["info"] = [int(info_sh), int(info_rep)]
        res["iters"] = [len(r_sh), len(r_rep)]
        res["x_rel"] = float((x_sh - x_rep).abs().max() / x_rep.abs().max())
        res["identical_on_ranks"] = same_on_ranks(x_sh)
        resid = b - A._apply(x_sh)
        res["relres"] = float(torch.linalg.norm(resid) / torch.linalg.norm(b))
        # the state machine started from the REPLICATED right-hand side (every rank computes p = M_BD b itself, no
        # exchange: what bench.py's timed loop does) gives the same iterates as the start from per-rank slices
        from cosmomap2_b200.pcg import make_solver, _run_loop
        s = make_solver(A, Mbd, b.numel())
        res["local_start_used_sharded"] = isinstance(s, distributed.ShardedPCG)
        s.start(b, None, 0.0, 1e-10)
        r_loc = []
        info_loc = _run_loop(s, 200, r_loc)
        x_loc = s.gather_x()
        res["local_start"] = [int(info_loc), len(r_loc), float((x_loc - x_sh).abs().max() / x_sh.abs().max()),
                              float(np.max(np.abs(np.array(r_loc) - np.array(r_sh))) / r_sh[0]) if len(r_loc) == len(r_sh) else 1.0,
                              bool(same_on_ranks(x_loc)), int(s.failed())]
        A.close()

    elif case == "m2_sharded":
        sc, A_loc, Mbd, b, npix = _rank_problem(cm, distributed, synthetic, rank, world, pol=3)
        n, r = 3 * npix, 8
        A = distributed.AllReduceLO(A_loc)
        g = torch.Generator(device="cuda")
        g.manual_seed(3)
        Zt = torch.randn((r, n), dtype=torch.float64, device="cuda", generator=g) / np.sqrt(n)
        AZt = torch.stack([A._apply(Zt[i]) for i in range(r)])
        Zd, AZd = cm.DeflationLO(Zt.t()), cm.DeflationLO(AZt.t())
        E = cm.CoarseLO(Zt.t(), AZt.t(), r, apply="eig")
        M2 = cm.TwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
        M2s = distributed.ShardedTwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
        v = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        y0, y1 = M2._apply(v), M2s._apply(v)
        res["rel_err"] = float((y1 - y0).abs().max() / y0.abs().max())
        res["identical_on_ranks"] = same_on_ranks(y1)
        A.close()

    elif case == "recover":
        # rank 1 arrives long after the timeout: the peer exchange fails on rank 0, every rank agrees,
        # switches to NCCL and the solve still returns the right answer
        os.environ["CM2_P2P_TIMEOUT"] = "0.3"
        sc, A_loc, Mbd, b, npix = _rank_problem(cm, distributed, synthetic, rank, world, pol=3)
        A = distributed.AllReduceLO(A_loc)
        had_p2p = A._p2p is not None
        # build the sharded solver now: its set-up is collective (IPC handle exchange) and would simply
        # make rank 0 wait for the late rank on the host; the timeout under test is the one in the kernel
        assert A.sharded_solver(Mbd) is not None or not had_p2p
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 1:
            time.sleep(2.0)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, info = cm.cg(A, b, M=Mbd, rtol=1e-12, maxiter=60)
        res["had_p2p"] = had_p2p
        res["fell_back"] = A._p2p is None
        res["info"] = int(info)
        resid = b - A._apply(x)
        res["relres"] = float(torch.linalg.norm(resid) / torch.linalg.norm(b))
        res["identical_on_ranks"] = same_on_ranks(x)
        # the plain all-reduce path recovers the same way
        A2 = distributed.AllReduceLO(A_loc)
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            time.sleep(2.0)
        os.environ["CM2_SHARDED_PCG"] = "0"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x2, info2 = cm.cg(A2, b, M=Mbd, rtol=1e-12, maxiter=60)
        os.environ["CM2_SHARDED_PCG"] = "1"
        res["fell_back_allreduce"] = A2._p2p is None
        res["x2_rel"] = float((x2 - x).abs().max() / x.abs().max())
        A2.close()

    else:
        raise SystemExit("unknown case %r" % case)

    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(allres, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    worker(sys.argv[1], sys.argv[2])
