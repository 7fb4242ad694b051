"""World-size-2 gloo (CPU) tests of the host-side sharding logic used by the multi-GPU path:
detector partition, map-domain all-reduce of sum_g P_g^T N_g P_g, common pixel set from summed
moments, and a sharded PCG solve that matches the single-process solve.  The per-rank operators
here are the NumPy oracle (the CUDA kernels need a GPU); the collective plumbing is the product's
(cosmomap2_b200/distributed.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse.linalg as spla
        import oracle
        from cosmomap2_b200 import distributed, synthetic
        sc = synthetic.raster_scan(48000, nside=32, ndet=6, nx=40, ny=24, samples_per_pixel=5.0, seed=11)
        pol = 3
        # ---- single-process reference solve (every rank computes it) ----
        pix = sc.pix.astype(np.int64)
        N = oracle.BlockLO(sc.ns, sc.weights)
        pts = oracle.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
        npix = pts.get_new_pixel[0]
        P = oracle.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        Mbd = oracle.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
        A = P.T * N * P
        b = P.T * (N * sc.d)
        x_ref, info = spla.cg(A, b, M=Mbd, rtol=1e-10, maxiter=50)
        assert info == 0
        # ---- sharded: this rank's detectors only ----
        (pix_l, phi_l, d_l), (lo, hi) = distributed.shard_tod([sc.pix.astype(np.int64), sc.phi, sc.d], sc.ndet,
                                                              sc.ns, world, rank)
        assert (hi - lo) == sc.ndet // world
        Nl = oracle.BlockLO(sc.ns, sc.weights[lo:hi])
        # common pixel set: moments summed over ranks, then the same mask everywhere
        ptl = oracle.ProcessTimeSamples(pix_l.copy(), sc.npix_full, pol=pol, phi=phi_l, w=Nl.diag,
                                        threshold_cond=1e30)      # local pass only to get the moments
        c, s = np.cos(2 * phi_l), np.sin(2 * phi_l)
        from oracle import cloops
        mom = np.stack(cloops.moments(pix_l, np.asarray(Nl.diag), c, s, pol, sc.npix_full))
        mom_t = torch.from_numpy(mom)
        distributed.all_reduce_sum_(mom_t)
        counts, cosine, sine, cos2, sin2, sincos = mom_t.numpy()
        det = cos2 * sin2 - sincos ** 2
        tr = cos2 + sin2
        with np.errstate(all="ignore"):
            sq = np.sqrt(tr * tr / 4 - det)
            cond = np.abs((tr / 2 + sq) / (tr / 2 - sq))
        good = (cond <= 1e3) & (counts > 2)
        assert int(good.sum()) == npix                     # same pixel set as the global run
        old2new = np.full(sc.npix_full, -1)
        old2new[good] = np.arange(npix)
        pl = pix_l.copy()
        pl[pl >= 0] = old2new[pl[pl >= 0]]
        ang = type("Ang", (), {"cos": c, "sin": s})
        Pl = oracle.SparseLO(npix, len(pl), pl, pol=pol, angle_processed=ang)
        A_local = Pl.T * Nl * Pl
        Ash = distributed.HostAllReduceLO(lambda v: A_local * v, pol * npix)
        bt = torch.from_numpy(Pl.T * (Nl * d_l))
        distributed.all_reduce_sum_(bt)
        b_sh = bt.numpy()
        assert np.allclose(b_sh, b, rtol=1e-12, atol=1e-12 * np.abs(b).max())
        v = np.random.default_rng(0).standard_normal(pol * npix)
        assert np.allclose(Ash * v, A * v, rtol=1e-12, atol=1e-12 * np.abs(A * v).max())
        x_sh, info = spla.cg(Ash, b_sh, M=Mbd, rtol=1e-10, maxiter=50)
        assert info == 0
        assert np.allclose(x_sh, x_ref, rtol=1e-8, atol=1e-10 * np.abs(x_ref).max())
        # every rank ends with bit-identical x (replicated CG scalars need no communication)
        xt = torch.from_numpy(x_sh.copy())
        gathered = [torch.zeros_like(xt) for _ in range(world)]
        dist.all_gather(gathered, xt)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_shard_detectors_partition():
    sys.path.insert(0, ROOT)
    from cosmomap2_b200.distributed import shard_detectors
    for ndet in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_detectors(ndet, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == ndet
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_sharded_solve_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def _worker_files(rank, world, port, out, workdir):
    """File-driven sharding: every rank reads ITS detector pairs of two CES files (read_ces_shard) and
    the summed operator equals the one built from read_multiple_ces of the whole files."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse.linalg as spla
        import oracle
        from cosmomap2_b200 import distributed, synthetic
        from cosmomap2_b200 import IOfiles as io
        files = [os.path.join(workdir, "ces_%d.hdf5" % k) for k in range(2)]
        if rank == 0:
            scans = [synthetic.raster_scan(5 * 3000, nside=32, ndet=5, nx=40, ny=24, samples_per_pixel=5.0, seed=20 + k,
                                           flag_turnarounds=True) for k in range(2)]
            obspix = np.unique(np.concatenate([sc.pix[sc.pix >= 0] for sc in scans]))
            for sc, path in zip(scans, files):
                idx = np.where(sc.pix >= 0, np.searchsorted(obspix, sc.pix), -1)
                cut = lambda a: [a[b * sc.ns:(b + 1) * sc.ns] for b in range(sc.ndet)]  # noqa: E731
                io.write_ces_to_hdf5(path, obspix, cut(idx), cut(sc.phi), [np.zeros(sc.ns, dtype=np.int32)] * sc.ndet,
                                     sc.ns, sc.sub_len, sc.sub_start, sum_=cut(sc.d), weight_sum=sc.weights)
        dist.barrier()
        pol = 1
        # ---- whole files, one process ----
        d, w, phi, pixs, hp, ground, subs, tst, ns_l, nb_l = io.read_multiple_ces(files, pol)
        npix = len(hp)
        P = oracle.SparseLO(npix, len(d), pixs, pol=pol)
        F = oracle.FilterLO(len(d), [subs, tst], ns_l, nb_l, pixs)
        A = P.T * F * P
        b = P.T * (F * d)
        # ---- this rank's pairs of every file ----
        parts = [io.read_ces_shard(f, pol, rank, world) for f in files]
        d_l = np.concatenate([p[0] for p in parts])
        pix_l = np.concatenate([p[3] for p in parts])
        Pl = oracle.SparseLO(npix, len(d_l), pix_l, pol=pol)
        Fl = oracle.FilterLO(len(d_l), [[p[8][0] for p in parts], [p[8][1] for p in parts]], [p[6] for p in parts],
                             [p[7] for p in parts], pix_l)
        assert sum(p[7] for p in parts) in (4, 6) and len(d_l) < len(d)          # 2+2 or 3+3 of the 5+5 pairs
        A_local = Pl.T * Fl * Pl
        Ash = distributed.HostAllReduceLO(lambda v: A_local * v, npix)
        bt = torch.from_numpy(np.ascontiguousarray(Pl.T * (Fl * d_l)))
        distributed.all_reduce_sum_(bt)
        assert np.allclose(bt.numpy(), b, rtol=1e-12, atol=1e-12 * np.abs(b).max())
        v = np.random.default_rng(1).standard_normal(npix)
        Av = A * v
        assert np.allclose(Ash * v, Av, rtol=1e-12, atol=1e-12 * np.abs(Av).max())
        hits = torch.from_numpy(np.bincount(pix_l[pix_l >= 0], minlength=npix).astype(np.float64))
        distributed.all_reduce_sum_(hits)
        assert np.array_equal(hits.numpy(), np.bincount(pixs[pixs >= 0], minlength=npix))
        Minv = oracle.lp.DiagonalOperator(np.where(hits.numpy() > 0, 1.0 / np.maximum(hits.numpy(), 1), 0.0))
        x1, _ = spla.cg(A, b, M=Minv, rtol=1e-12, maxiter=15)
        x2, _ = spla.cg(Ash, bt.numpy(), M=Minv, rtol=1e-12, maxiter=15)
        assert np.allclose(A * x2, A * x1, rtol=1e-8, atol=1e-9 * np.abs(b).max())
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_file_driven_sharded_operator_world_size_2(tmp_path):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_files, args=(r, 2, port, out, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
