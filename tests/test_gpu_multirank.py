"""The multi-GPU path on real GPUs.

* one GPU: the pixel-sharded PCG kernel (cm2_pcg_bd_sharded: reduce-scatter by peer loads + M_BD + CG vector
  work + all-gather by peer stores) with all ranks emulated in ONE cooperative launch, against the
  single-GPU solver and SciPy's cg over the oracle;
* >= 2 GPUs: two real ranks under torchrun (NCCL + NVLink peer memory): the peer-memory all-reduce equals the
  rank-ordered sum bit for bit and NCCL to rounding; the sharded solve equals the replicated solve, the
  one-process solve and the oracle; the sharded two-level preconditioner equals the replicated one; a rank
  that arrives after the timeout makes every rank fall back to NCCL and the answer is still right.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scan(nt=120000, ndet=8, seed=2):
    from cosmomap2_b200 import synthetic
    return synthetic.raster_scan(nt, nside=64, ndet=ndet, nx=70, ny=40, samples_per_pixel=5.0, seed=seed,
                                 flag_turnarounds=True)


@pytest.mark.parametrize("world,pol,correlated", [(1, 3, False), (2, 3, True), (2, 1, True), (3, 2, True), (4, 3, False),
                                                  (4, 3, True), (8, 3, True), (8, 1, False)])
def test_sharded_pcg_emulated_equals_single_gpu_solver(world, pol, correlated):
    import torch
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import _device as dv
    from multirank import EmulatedShardedPCG, split_problem
    sc = _scan()
    A, A_locals, Mbd, b, npix = split_problem(sc, pol, world, cm, correlated=correlated)
    bd = dv.to_dev_f64(b)
    # reference: the single-GPU solver, a fixed number of iterations with the exit test off
    res = []
    x_ref, info = cm.cg(A, bd, M=Mbd, rtol=0.0, atol=0.0, maxiter=6, residuals=res)
    em = EmulatedShardedPCG(A_locals, Mbd)
    em.start(bd, atol=0.0, rtol=0.0)
    hist = []
    for it in range(6):
        sc_all = em.scalars()
        hist.append(np.sqrt(sc_all[0][3]))
        for g in range(1, world):                     # bit-identical scalars on every rank
            assert np.array_equal(sc_all[g][:9], sc_all[0][:9])
        em.step()
        for g in range(1, world):                     # ... and bit-identical search directions
            assert torch.equal(em.p[g], em.p[0])
    sc_all = em.scalars()
    assert all(s[9] == 0.0 for s in sc_all), "a flag wait timed out"
    assert int(sc_all[0][8]) == 6
    x = em.solution()
    bn = float(torch.linalg.norm(bd))
    assert float((x - x_ref).abs().max() / x_ref.abs().max()) < 1e-11
    assert np.max(np.abs(np.array(hist) - np.array(res[:6]))) / bn < 1e-12


def test_sharded_pcg_emulated_stops_like_scipy():
    """Exit rule and iteration count: the emulated sharded solver against SciPy's cg over the oracle."""
    import scipy.sparse.linalg as spla
    import torch
    import cosmomap2_b200 as cm
    import oracle
    from cosmomap2_b200 import _device as dv
    from multirank import EmulatedShardedPCG, split_problem
    pol, world = 3, 4
    sc = _scan(seed=4)
    from cosmomap2_b200 import synthetic
    A, A_locals, Mbd, b, npix = split_problem(sc, pol, world, cm, correlated=True)
    pix = sc.pix.astype(np.int64)
    Nw = oracle.BlockLO(sc.ns, sc.weights)
    No = oracle.BlockLO(sc.ns, synthetic.toeplitz_bands(sc.ndet, 3, seed=9), offdiag=True)
    pts = oracle.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=Nw.diag)
    Po = oracle.SparseLO(pts.get_new_pixel[0], sc.nt, pix, pol=pol, angle_processed=pts)
    Mo = oracle.BlockDiagonalPreconditionerLO(pts, pts.get_new_pixel[0], pol=pol)
    Ao, bo = Po.T * No * Po, Po.T * (No * sc.d)
    count = [0]
    xo, info = spla.cg(Ao, bo, M=Mo, rtol=1e-10, maxiter=50, callback=lambda xk: count.__setitem__(0, count[0] + 1))
    assert info == 0
    em = EmulatedShardedPCG(A_locals, Mbd)
    em.start(dv.to_dev_f64(b), atol=0.0, rtol=1e-10)
    iters = 0
    while em.scalars()[0][7] == 0.0 and iters < 50:
        em.step()
        iters += 1
    assert iters == count[0]
    em.step()                                          # a queued extra iteration is a no-op once `done` is set
    assert int(em.scalars()[0][8]) == iters
    x = dv.to_host(em.solution())
    assert np.max(np.abs(x - xo)) / np.max(np.abs(xo)) < 1e-10


def test_sharded_pcg_emulated_zero_rhs_and_empty_slices():
    import torch
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import _device as dv
    from multirank import EmulatedShardedPCG, split_problem
    sc = _scan(nt=40000, ndet=8)
    A, A_locals, Mbd, b, npix = split_problem(sc, 3, 8, cm)
    em = EmulatedShardedPCG(A_locals, Mbd)
    em.start(dv.zeros_f64(3 * npix), atol=0.0, rtol=1e-8)
    s = em.scalars()[0]
    assert s[7] == 1.0 and s[9] == 0.0                # b = 0: done at once, x = 0
    assert float(em.solution().abs().max()) == 0.0


def test_cooperative_launch_refused_falls_back():
    """a refused cooperative launch must degrade to the 3-kernel tail, not abort cg()."""
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import _device as dv
    from multirank import split_problem
    sc = _scan()
    A, _, Mbd, b, npix = split_problem(sc, 3, 1, cm, correlated=True)
    r0, r1 = [], []
    x0, info0 = cm.cg(A, b, M=Mbd, rtol=1e-10, maxiter=60, residuals=r0)
    assert info0 == 0, r0
    old = dv.call("cm2_pcg_bd_iter_refuse", 1)
    try:
        x1, info1 = cm.cg(A, b, M=Mbd, rtol=1e-10, maxiter=60, residuals=r1)
    finally:
        dv.call("cm2_pcg_bd_iter_refuse", old)
    assert info1 == 0, (r0, r1)
    assert len(r0) == len(r1)
    assert np.max(np.abs(x1 - x0)) / np.max(np.abs(x0)) < 1e-12


# ---- real ranks -------------------------------------------------------------------------------------
def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run_ranks(case, nproc, tmp_path, timeout=600):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = os.path.join(str(tmp_path), "%s.json" % case)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multirank.py"), case, out]
    env = dict(os.environ)
    env.pop("CM2_SHARDED_PCG", None)
    env.pop("CM2_P2P_ALLREDUCE", None)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env)
    assert p.returncode == 0, p.stdout[-4000:]
    return json.load(open(out))


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs (ranks that wait on each other must not share a GPU)")


@needs2
def test_ranks_p2p_allreduce(tmp_path):
    for r in _run_ranks("p2p_allreduce", 2, tmp_path):
        assert r["ok"]


@needs2
@pytest.mark.parametrize("case", ["sharded_solve", "sharded_solve_pol1"])
def test_ranks_sharded_solve(case, tmp_path):
    allr = _run_ranks(case, 2, tmp_path)
    for r in allr:
        assert r["sharded_available"] and r["used_sharded"]
        assert r["info"] == [0, 0]
        assert r["iters"][0] == r["iters"][1] == r["iters_single"]
        assert r["identical_on_ranks"] and r["slice_equal"] and r["npix_equal"]
        assert r["x_rel"] < 1e-10 and r["res_rel"] < 1e-10 and r["x_rel_single"] < 1e-10
    assert allr[0]["x_rel_oracle"] < 1e-10
    assert abs(allr[0]["iters_oracle"] - allr[0]["iters"][0]) <= 1


@needs2
def test_ranks_sharded_solve_many_iterations(tmp_path):
    for r in _run_ranks("sharded_solve_toeplitz", 2, tmp_path):
        assert r["used_sharded"] and r["info"] == [0, 0]
        assert abs(r["iters"][0] - r["iters"][1]) <= 1
        assert r["x_rel"] < 1e-8 and r["relres"] < 2e-10 and r["identical_on_ranks"]
        # the exchange-free start from the replicated right-hand side: same exit, same number of residuals, same
        # residual history (relative to ||b||-scaled first residual) and solution, bit-identical on the ranks
        info_loc, n_loc, x_rel_loc, hist_rel, same, failed = r["local_start"]
        assert r["local_start_used_sharded"] and info_loc == 0 and n_loc == r["iters"][0] and failed == 0
        assert x_rel_loc < 1e-10 and hist_rel < 1e-10 and same


@needs2
def test_ranks_sharded_two_level_preconditioner(tmp_path):
    for r in _run_ranks("m2_sharded", 2, tmp_path):
        assert r["rel_err"] < 1e-11 and r["identical_on_ranks"]


@needs2
def test_ranks_recover_from_a_late_rank(tmp_path):
    for r in _run_ranks("recover", 2, tmp_path):
        assert r["had_p2p"] and r["fell_back"] and r["fell_back_allreduce"]
        assert r["info"] == 0 and r["relres"] < 1e-10 and r["identical_on_ranks"] and r["x2_rel"] < 1e-10
