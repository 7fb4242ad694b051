#!/usr/bin/env python
"""The reference's real-data recipe (src/test_M2_precond_onto_real_data.py:54-122) end to end from
files: CES files in the AnalysisBackend HDF5 schema -> read_multiple_ces -> ProcessTimeSamples ->
SparseLO / FilterLO / M_BD -> PCG.  The real Polarbear files are not public, so the files are first
written (write_ces_to_hdf5) from a synthetic raster scan; everything after that line is the reference
script with `from cosmomap2_b200 import *` in place of `from interfaces import *`.

    python examples/solve_from_ces_files.py [--nces 2] [--npair 16] [--ns 200000] [--poly-order 0]
    torchrun --nproc-per-node N ... : every rank reads only its detector pairs (read_ces_shard)
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nces", type=int, default=2)
    ap.add_argument("--npair", type=int, default=16)
    ap.add_argument("--ns", type=int, default=200000, help="samples per detector pair and CES")
    ap.add_argument("--nside", type=int, default=256)
    ap.add_argument("--poly-order", type=int, default=0)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--dir", default=None, help="where the CES files are written (default: a temp dir)")
    args = ap.parse_args()

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from cosmomap2_b200 import (ProcessTimeSamples, SparseLO, FilterLO, BlockDiagonalPreconditionerLO, cg,
                                read_multiple_ces, read_ces_shard, write_ces_to_hdf5, reorganize_map, distributed,
                                synthetic)

    pol, out = 3, {"world": world}
    workdir = args.dir or tempfile.mkdtemp(prefix="cm2_ces_")
    files = [os.path.join(workdir, "ces_%d.hdf5" % k) for k in range(args.nces)]
    if rank == 0:                                          # ---- stand-in for the AnalysisBackend ----
        t0 = time.perf_counter()
        nbytes = 0
        scans = [synthetic.raster_scan(args.npair * args.ns, nside=args.nside, ndet=args.npair, nx=300, ny=150,
                                       samples_per_pixel=8.0, seed=k, flag_turnarounds=True) for k in range(args.nces)]
        # one observed-pixel list for all CES (the reference's common_obspix), `pixel` indexes into it
        obspix = np.unique(np.concatenate([sc.pix[sc.pix >= 0] for sc in scans]))
        for sc, path in zip(scans, files):
            idx = np.where(sc.pix >= 0, np.searchsorted(obspix, sc.pix), -1)
            cut = lambda a: [a[b * sc.ns:(b + 1) * sc.ns] for b in range(sc.ndet)]  # noqa: E731
            ground = ((np.arange(sc.ns) // 40) % 200).astype(np.int32)
            write_ces_to_hdf5(path, obspix, cut(idx), cut(sc.phi), [ground] * sc.ndet, sc.ns, sc.sub_len,
                              sc.sub_start, dif=cut(sc.d), weight_dif=sc.weights, sum_=cut(sc.d), weight_sum=sc.weights)
            nbytes += os.path.getsize(path)
        del scans
        out["files"] = dict(n=len(files), GB=nbytes / 1e9, write_s=time.perf_counter() - t0)
    if world > 1:
        dist.barrier()

    # ---- the reference script from here on ------------------------------------------------------
    t0 = time.perf_counter()
    if world == 1:
        d, weight, polang, pixs, hp_pixs, ground, subscan_nsample, tstart, samples_per_bolopair, bolos_per_ces = \
            read_multiple_ces(files, pol)
    else:                                                  # same tuple, this rank's detector pairs only
        parts = [read_ces_shard(f, pol, rank, world) for f in files]
        d, weight, polang, pixs, ground = [np.concatenate([np.atleast_1d(p[k]) for p in parts]) for k in (0, 1, 2, 3, 5)]
        hp_pixs = parts[-1][4]
        subscan_nsample, tstart = [p[8][0] for p in parts], [p[8][1] for p in parts]
        samples_per_bolopair, bolos_per_ces = [p[6] for p in parts], [p[7] for p in parts]
    out["read"] = dict(seconds=time.perf_counter() - t0, samples=int(len(d)),
                       GBps=(len(d) * 28 / 1e9) / (time.perf_counter() - t0))
    nt = len(d)
    npix = len(hp_pixs)                                    # src/test_M2_precond_onto_real_data.py:70-73
    pts = ProcessTimeSamples(pixs, npix, obspix=hp_pixs, pol=pol, phi=polang, ground=ground,
                             comm=(True if world > 1 else None))
    npix, obspix = pts.get_new_pixel
    P = SparseLO(npix, nt, pixs, pol=pol, angle_processed=pts)
    F = FilterLO(nt, [subscan_nsample, tstart], samples_per_bolopair, bolos_per_ces, P.pairs,
                 poly_order=args.poly_order)
    Mbd = BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A_local = P.T * F * P
    A = distributed.AllReduceLO(A_local) if world > 1 else A_local
    b = P.T * (F * d)
    if world > 1:
        bt = torch.from_numpy(b).cuda()
        distributed.all_reduce_sum_(bt)
        b = bt.cpu().numpy()
    res = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = cg(A, b, M=Mbd, tol=args.rtol, maxiter=2000, residuals=res)
    torch.cuda.synchronize()
    out["solve"] = dict(info=int(info), iterations=len(res) - 1, seconds=time.perf_counter() - t0, npix=int(npix),
                        nt=int(nt), relres=float(np.linalg.norm(b - A * x) / np.linalg.norm(b)))
    hp = reorganize_map(x, obspix, npix, args.nside, pol)
    out["map"] = dict(observed=int(np.count_nonzero(hp[0])), rms_I=float(np.std(hp[0][obspix])))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        if hasattr(A, "close"):
            A.close()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
