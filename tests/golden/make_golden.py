#!/usr/bin/env python
"""
Generate the golden vectors under tests/golden/ by running the reference's OWN code
(/root/reference, through oracle/refrun.py: lexical py2->py3 + a weave.inline shim that compiles
the reference's C++ loop bodies with g++).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are seeded and stored next to the outputs, so the parity tests never depend on a RNG
stream.  The SciPy PCG (scipy.sparse.linalg.cg, SciPy %(scipy)s on this box) drives the
reference operators exactly as src/test_BD_precond_onto_real_data.py:47 and
tests/test_2level_preconditioner.py:52 do.
"""
import os
import sys

import numpy as np
import scipy
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refrun  # noqa: E402

R = refrun.load_reference()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f kB)" % (path, os.path.getsize(path) / 1e3))


def make_scan(rng, nt=2400, npix_main=60, npix_old=70, flag_frac=0.03):
    """Random pointing with flags plus a few pathological pixels (few hits / degenerate angles)."""
    pix = rng.integers(0, npix_main, size=nt).astype(np.int64)
    phi = R.angles_gen(float(rng.uniform(0, np.pi)), nt)
    k = nt - 20
    pix[k] = 60                               # 1 hit
    pix[k + 1:k + 3] = 61                     # 2 hits
    pix[k + 3:k + 6] = 62                     # 3 hits, identical angle -> singular 2x2 block
    phi[k + 3:k + 6] = phi[k + 3]
    pix[k + 6:k + 9] = 63                     # 3 hits, distinct angles
    pix[k + 9:k + 15] = 64                    # 6 hits, nearly identical angles -> cond > 1e3
    phi[k + 9:k + 15] = phi[k + 9] + 1e-4 * np.arange(6)
    flags = rng.random(nt) < flag_frac
    flags[k:] = False
    pix[flags] = -1
    return pix, phi, npix_old


def case_process_and_pointing():
    for pol in (1, 2, 3):
        for weighted in (False, True):
            rng = np.random.default_rng(100 + 10 * pol + int(weighted))
            pix0, phi, npix_old = make_scan(rng)
            nt = len(pix0)
            nb = 6
            tdiag = rng.random(nb) + 0.5
            w = None
            if weighted:
                N = R.BlockLO(nt // nb, tdiag, offdiag=False)
                w = N.diag.copy()
            pix = pix0.copy()
            pts = R.ProcessTimeSamples(pix, npix_old, pol=pol, phi=phi, w=w)
            npix, obspix = pts.get_new_pixel
            P = R.SparseLO(npix, nt, pix, pol=pol, angle_processed=pts)
            x = rng.standard_normal(pol * npix)
            d = rng.standard_normal(nt)
            Mbd = R.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
            Bd = R.BlockDiagonalLO(pts, npix, pol=pol)
            out = dict(pix_in=pix0, phi=phi, npix_old=npix_old, pol=pol, tdiag=tdiag,
                       weighted=weighted, w=(w if w is not None else np.zeros(0)),
                       pix_out=pix, npix_new=npix, obspix=np.asarray(obspix),
                       old2new=np.asarray(pts.old2new), mask=np.asarray(pts.mask),
                       x=x, d=d, Px=P * x, Ptd=P.T * d, PtPx=P.T * (P * x),
                       Mbd_x=Mbd * x, Bd_x=Bd * x)
            for nm in ("counts", "cosine", "sine", "cos2", "sin2", "sincos", "cos", "sin"):
                if hasattr(pts, nm):
                    out["pts_" + nm] = np.asarray(getattr(pts, nm))
            if weighted:
                out["PtNPx"] = P.T * (N * (P * x))
            save("pointing_pol%d_%s" % (pol, "w" if weighted else "u"), **out)


def case_obspix2():
    """The SetObspix / compute_arrays path (process_ces.py:80-83, 94-189)."""
    for pol in (1, 3):
        rng = np.random.default_rng(300 + pol)
        nt, npix_old = 1500, 40
        pix0 = rng.integers(0, npix_old, size=nt).astype(np.int64)
        pix0[rng.random(nt) < 0.05] = -1
        phi = R.angles_gen(0.3, nt)
        obspix = np.arange(1000, 1000 + npix_old, dtype=np.int64)
        obspix2 = obspix[3:31].copy()
        pix = pix0.copy()
        pts = R.ProcessTimeSamples(pix, npix_old, obspix=obspix.copy(), pol=pol, phi=phi,
                                   obspix2=obspix2)
        npix, op = pts.get_new_pixel
        out = dict(pix_in=pix0, phi=phi, npix_old=npix_old, pol=pol, obspix=obspix,
                   obspix2=obspix2, pix_out=pix, npix_new=npix, obspix_out=np.asarray(op),
                   old2new=np.asarray(pts.old2new))
        for nm in ("counts", "cosine", "sine", "cos2", "sin2", "sincos"):
            if hasattr(pts, nm):
                out["pts_" + nm] = np.asarray(getattr(pts, nm))
        save("obspix2_pol%d" % pol, **out)


def case_toeplitz():
    rng = np.random.default_rng(7)
    out = {}
    n = 257
    v = rng.standard_normal(n)
    out["v"] = v
    for L in (1, 2, 5, 17, 64):
        a = rng.random(L)
        a[0] += 2.0
        T = R.ToeplitzLO(a, n)
        out["a%d" % L] = a
        out["y%d" % L] = T * v
    nb, bs = 4, 300
    t = [rng.random(3) + np.array([2., 0, 0]) for _ in range(nb)]
    N = R.BlockLO(bs, t, offdiag=True)
    vv = rng.standard_normal(nb * bs)
    out["blk_t"] = np.array(t)
    out["blk_v"] = vv
    out["blk_y"] = N * vv
    out["blk_diag"] = np.asarray(N.diag)
    tw = rng.random(nb) + 0.5
    Nw = R.BlockLO(bs, tw, offdiag=False)
    out["white_t"] = tw
    out["white_y"] = Nw * vv
    out["white_diag"] = np.asarray(Nw.diag)
    W = R.WeightingLO([2, 3], [100, 200], rng.random(5))
    dd = rng.standard_normal(W.size)
    out["wt_weights"] = np.asarray(W.weights)
    out["wt_d"] = dd
    out["wt_y"] = W * dd.copy()
    save("noise_ops", **out)


def make_subscans(rng, ns, nsub, gap=7):
    """Subscan (lengths, starts) inside [0, ns) with gaps; trailing samples left outside."""
    edges = np.sort(rng.choice(np.arange(gap, ns - gap), size=nsub - 1, replace=False))
    starts = np.concatenate([[3], edges + gap])
    ends = np.concatenate([edges, [ns - 5]])
    keep = ends > starts
    return (ends - starts)[keep].astype(np.int64), starts[keep].astype(np.int64)


def case_filter():
    rng = np.random.default_rng(11)
    nsamples = [400, 300]
    nbolos = [3, 2]
    subs, tst = [], []
    for ns in nsamples:
        L, S = make_subscans(rng, ns, 6)
        subs.append(L)
        tst.append(S)
    nt = sum(a * b for a, b in zip(nsamples, nbolos))
    npix = 30
    pix = rng.integers(0, npix, size=nt).astype(np.int64)
    pix[rng.random(nt) < 0.1] = -1
    # one fully flagged subscan: detector 1 of CES 0, subscan 2
    s0 = nsamples[0] * 1 + tst[0][2]
    pix[s0:s0 + subs[0][2]] = -1
    d = rng.standard_normal(nt) + 3.0
    F = R.FilterLO(nt, [subs, tst], nsamples, nbolos, pix)
    out = dict(nsamples=np.array(nsamples), nbolos=np.array(nbolos), pix=pix, d=d, Fd=F * d, npix=npix)
    for i in range(2):
        out["sub_len%d" % i] = subs[i]
        out["sub_start%d" % i] = tst[i]
    # single-CES scalar form of the ctor (linearoperators.py:269-273)
    F1 = R.FilterLO(nsamples[0] * nbolos[0], [subs[0], tst[0]], nsamples[0], nbolos[0],
                    pix[:nsamples[0] * nbolos[0]])
    out["Fd_single"] = F1 * d[:nsamples[0] * nbolos[0]]
    # A = P^T F P through the reference operators, pol 1 and 3
    for pol in (1, 3):
        phi = R.angles_gen(0.7, nt)
        pp = pix.copy()
        pts = R.ProcessTimeSamples(pp, npix, pol=pol, phi=phi)
        npn = pts.get_new_pixel[0]
        P = R.SparseLO(npn, nt, pp, pol=pol, angle_processed=pts)
        Fp = R.FilterLO(nt, [subs, tst], nsamples, nbolos, P.pairs)
        x = rng.standard_normal(pol * npn)
        out["phi"] = phi
        out["A_x_pol%d" % pol] = x
        out["A_y_pol%d" % pol] = P.T * (Fp * (P * x))
        out["A_pix_pol%d" % pol] = pp
        out["A_npix_pol%d" % pol] = npn
    save("filter_ops", **out)


def case_pcg_and_deflation():
    """PCG with Toeplitz noise + M_BD; ARPACK deflation space; CoarseLO / DeflationLO / M2;
    in-tree arnoldi (tests/test_2level_preconditioner.py, tests/test_coarse_operator.py)."""
    for pol in (1, 2, 3):
        rng = np.random.default_rng(500 + pol)
        nt, npix_old, nb = 6000, 48, 4
        pix0 = rng.integers(0, npix_old, size=nt).astype(np.int64)
        pix0[rng.random(nt) < 0.02] = -1
        phi = R.angles_gen(float(rng.uniform(0, np.pi)), nt)
        d = rng.random(nt)
        t = [np.array([1.0 + rng.random(), -0.3 * rng.random(), 0.1 * rng.random()]) for _ in range(nb)]
        N = R.BlockLO(nt // nb, t, offdiag=True)
        pix = pix0.copy()
        pts = R.ProcessTimeSamples(pix, npix_old, pol=pol, phi=phi)
        npix = pts.get_new_pixel[0]
        P = R.SparseLO(npix, nt, pix, pol=pol, angle_processed=pts)
        Mbd = R.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
        B = R.BlockDiagonalLO(pts, npix, pol=pol)
        A = P.T * N * P
        b = P.T * N * d
        n = pol * npix
        # --- PCG, M_BD
        hist = []
        xs = []

        def cb(xk):
            xs.append(xk.copy())
            hist.append(np.linalg.norm(b - A * xk))
        x, info = spla.cg(A, b, x0=np.zeros(n), M=Mbd, rtol=1e-8, maxiter=200, callback=cb)
        out = dict(pix_in=pix0, phi=phi, d=d, t=np.array(t), nb=nb, npix_old=npix_old, pol=pol,
                   npix=npix, b=b, cg_x=x, cg_info=info, cg_iters=len(hist), cg_hist=np.array(hist))
        # iterates after fixed iteration counts (exit test off) and the fully converged solution: the
        # quantities north_star's 1e-10 applies to (a solution stopped at rtol = 1e-8 is only defined
        # to about 1e-8 * cond, whatever the implementation)
        xs_fix = []
        spla.cg(A, b, x0=np.zeros(n), M=Mbd, rtol=0.0, atol=0.0, maxiter=8, callback=lambda xk: xs_fix.append(xk.copy()))
        out["cg_x_iter"] = np.array(xs_fix)
        hist_t = []
        xt, info_t = spla.cg(A, b, x0=np.zeros(n), M=Mbd, rtol=1e-13, maxiter=500,
                             callback=lambda xk: hist_t.append(np.linalg.norm(b - A * xk)))
        out.update(cg_x_tight=xt, cg_info_tight=info_t, cg_iters_tight=len(hist_t), cg_hist_tight=np.array(hist_t))
        # --- deflation space through ARPACK exactly as tests/test_2level_preconditioner.py:33
        x0 = np.ones(n)
        eigv, Z = spla.eigsh(A, M=B, Minv=Mbd, k=5, v0=x0, which="SM", ncv=15, tol=1e-10)
        r = Z.shape[1]
        Az = Z * 0.
        for i in range(r):
            Az[:, i] = A * Z[:, i]
        E_lu = R.CoarseLO(Z, Az, r)
        E_eig = R.CoarseLO(Z, Az, r, apply="eig")
        Zd = R.DeflationLO(Z)
        AZd = R.DeflationLO(Az)
        v_r = rng.standard_normal(r)
        v_n = rng.standard_normal(n)
        I = R.lp.IdentityOperator(n)
        Rop = I - AZd * E_eig * Zd.T
        M2 = Mbd * Rop + Zd * E_eig * Zd.T
        out.update(eigv=eigv, Z=Z, Az=Az, v_r=v_r, v_n=v_n, E=R.dgemm(Z, Az.T),
                   Elu_v=E_lu * v_r, Eeig_v=E_eig * v_r, Zd_v=Zd * v_r, ZdT_v=Zd.T * v_n,
                   M2_v=M2 * v_n, R_v=Rop * v_n)
        hist2 = []
        x2, info2 = spla.cg(A, b, x0=np.zeros(n), M=M2, rtol=1e-8, maxiter=200,
                            callback=lambda xk: hist2.append(np.linalg.norm(b - A * xk)))
        out.update(cg2_x=x2, cg2_info=info2, cg2_iters=len(hist2), cg2_hist=np.array(hist2))
        hist2t = []
        x2t, info2t = spla.cg(A, b, x0=np.zeros(n), M=M2, rtol=1e-13, maxiter=500,
                              callback=lambda xk: hist2t.append(np.linalg.norm(b - A * xk)))
        out.update(cg2_x_tight=x2t, cg2_info_tight=info2t, cg2_iters_tight=len(hist2t), cg2_hist_tight=np.array(hist2t),
                   E_eigvals=np.linalg.eigvalsh(0.5 * (E_lu.E + E_lu.E.T)) if hasattr(E_lu, "E") else
                   np.linalg.eigvalsh(0.5 * (R.dgemm(Z, Az.T) + R.dgemm(Z, Az.T).T)))
        # --- in-tree Arnoldi on Mbd*A (src/test_M2_precond_onto_real_data.py:42-44)
        # tol=1e-3 terminates through the reference's element-wise stop test (deflationlib.py:101);
        # tol=1e-5 runs into inner_m and raises (deflationlib.py:111-112) -- both are recorded.
        vs, hs, m = R.arnoldi(Mbd * A, Mbd * b, x0=np.ones(n), tol=1e-3, inner_m=n - 1)
        H = R.build_hess(hs, m)
        try:
            R.arnoldi(Mbd * A, Mbd * b, x0=np.ones(n), tol=1e-5, inner_m=n - 1)
            raised = 0
        except RuntimeError:
            raised = 1
        out.update(arn_m=m, arn_V=np.array(vs), arn_H=H, arn_raises_tol1em5=raised)
        save("solve_pol%d" % pol, **out)


def case_next_rows():
    """GroundFilterLO (linearoperators.py:24-61), the Legendre path of FilterLO
    (:170-204, serial polyfilter; the multiprocessing wrapper is Python-2 only) and reorganize_map
    (healpy_functions.py:50-105, restated without healpy in the oracle; not run here)."""
    rng = np.random.default_rng(21)
    out = {}
    nt = 3000
    ground = rng.integers(0, 25, size=nt).astype(np.int64)
    ground[rng.random(nt) < 0.07] = -1
    v = rng.standard_normal(nt)
    G = R.GroundFilterLO(ground.copy())
    out.update(ground=ground, gv=v, gFv=G * v, g_nbins=G.nbins)
    # Legendre filter: 2 CES, flags, one subscan with fewer unflagged samples than the order
    nsamples, nbolos = [400, 300], [2, 2]
    subs, tst = [], []
    for ns in nsamples:
        L, S = make_subscans(rng, ns, 5)
        subs.append(L)
        tst.append(S)
    ntl = sum(a * b for a, b in zip(nsamples, nbolos))
    pix = rng.integers(0, 30, size=ntl).astype(np.int64)
    pix[rng.random(ntl) < 0.15] = -1
    s0 = nsamples[0] * 1 + tst[0][1]
    pix[s0:s0 + subs[0][1]] = -1
    pix[s0] = 3                                          # 1 unflagged sample <= poly_order
    s1 = tst[0][2]
    pix[s1:s1 + subs[0][2]] = np.abs(pix[s1:s1 + subs[0][2]])   # a subscan without any flag
    d = rng.standard_normal(ntl) + np.linspace(0, 5, ntl)
    out.update(l_nsamples=np.array(nsamples), l_nbolos=np.array(nbolos), l_pix=pix, l_d=d)
    for i in range(2):
        out["l_sub_len%d" % i] = subs[i]
        out["l_sub_start%d" % i] = tst[i]
    for order in (1, 2, 3):
        F = R.FilterLO(ntl, [subs, tst], nsamples, nbolos, pix, poly_order=order, npool=1)
        out["l_Fd_order%d" % order] = F.polyfilter(d)
        F.procs.terminate()
    save("next_rows", **out)


def case_fused_chains():
    """The compositions the product collapses into single TOD passes, through the reference's own
    operators: A = P.T*N*P with N = BlockLO(bs, t, offdiag=True) and bands of 2..9 coefficients
    (tests/test_2level_preconditioner.py:16-29) on a scan with runs of equal pixels, flags and noise
    blocks that are not multiples of the kernel's 8-sample chunks; F*P and P.T*F*N*F*P with the
    offset filter (flagged samples inside subscans, one fully flagged subscan)."""
    rng = np.random.default_rng(31)
    out = {}
    nb, bs = 5, 403
    nt = nb * bs
    npix = 40
    # a scan: the pixel index drifts slowly, so consecutive samples share pixels (runs of 1..9 samples)
    steps = rng.random(nt) < 0.25
    pix = (np.cumsum(steps) % npix).astype(np.int64)
    pix[rng.random(nt) < 0.05] = -1
    phi = R.angles_gen(0.3, nt)
    out.update(pix=pix, phi=phi, npix=npix, nb=nb, bs=bs)
    # subscans: per block (detector) 4 subscans with gaps
    L, S = make_subscans(rng, bs, 4)
    out.update(sub_len=L, sub_start=S)
    for pol in (1, 2, 3):
        pp = pix.copy()
        pts = R.ProcessTimeSamples(pp, npix, pol=pol, phi=phi)
        npn = pts.get_new_pixel[0]
        P = R.SparseLO(npn, nt, pp, pol=pol, angle_processed=pts)
        x = rng.standard_normal(pol * npn)
        out["pix_pol%d" % pol] = pp
        out["npix_pol%d" % pol] = npn
        out["x_pol%d" % pol] = x
        for nband in (2, 3, 5, 9):
            t = [rng.standard_normal(nband) * 0.5 ** np.arange(nband) + np.eye(1, nband)[0] * 2.0 for _ in range(nb)]
            N = R.BlockLO(bs, t, offdiag=True)
            out["t_pol%d_nband%d" % (pol, nband)] = np.array(t)
            out["PtNPx_pol%d_nband%d" % (pol, nband)] = P.T * (N * (P * x))
        # offset filter: the same scan with one subscan of block 1 fully flagged, processed on its own
        pf = pix.copy()
        s0 = bs * 1 + S[2]
        pf[s0:s0 + L[2]] = -1
        ptsf = R.ProcessTimeSamples(pf, npix, pol=pol, phi=phi)
        npf = ptsf.get_new_pixel[0]
        Pf = R.SparseLO(npf, nt, pf, pol=pol, angle_processed=ptsf)
        F = R.FilterLO(nt, [L, S], bs, nb, Pf.pairs)
        t = [rng.standard_normal(4) * 0.5 ** np.arange(4) + np.eye(1, 4)[0] * 2.0 for _ in range(nb)]
        N = R.BlockLO(bs, t, offdiag=True)
        xf = rng.standard_normal(pol * npf)
        dF = F * (Pf * xf)
        out["pixf_pol%d" % pol] = pf
        out["npixf_pol%d" % pol] = npf
        out["xf_pol%d" % pol] = xf
        out["tF_pol%d" % pol] = np.array(t)
        out["FPx_pol%d" % pol] = dF
        out["PtFNFPx_pol%d" % pol] = Pf.T * (F * (N * dF))
    save("fused_chains", **out)


if __name__ == "__main__":
    print("scipy", scipy.__version__, "numpy", np.__version__)
    if len(sys.argv) > 1:                      # python make_golden.py case_pcg_and_deflation ...
        for nm in sys.argv[1:]:
            globals()[nm]()
        sys.exit(0)
    case_process_and_pointing()
    case_obspix2()
    case_toeplitz()
    case_filter()
    case_pcg_and_deflation()
    case_next_rows()
    case_fused_chains()
