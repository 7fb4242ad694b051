"""Small-scale analogues of BASELINE.json's configurations, GPU product vs oracle end to end:
identical seeded inputs, the same algorithm on both sides (SciPy cg over the oracle vs the device
PCG), compared on solution, exit code, iteration count (+-1) and residual history."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import golden_cases as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import cosmomap2_b200
    return cosmomap2_b200


def _build(impl, sc, pol, noise=None, filt=False, w_from_noise=True):
    pix = sc.pix.astype(np.int64)
    N = None
    if noise == "white":
        N = impl.BlockLO(sc.ns, sc.weights)
    elif noise is not None:
        N = impl.BlockLO(sc.ns, noise, offdiag=True)
    w = N.diag if (noise == "white" and w_from_noise) else None
    pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=w)
    npix = pts.get_new_pixel[0]
    P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
    Mbd = impl.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    if filt:
        F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix)
        A = P.T * F * P
        b = P.T * (F * sc.d)
    elif N is not None:
        A = P.T * N * P
        b = P.T * (N * sc.d)
    else:
        A = P.T * P
        b = P.T * sc.d
    return npix, P, Mbd, A, b, pts


TIGHT = 1e-13      # "fully converged": the solution is then defined to ~cond * 1e-13 whatever the rounding path


def _solve_both(cm, sc, pol, rtol, maxiter, tight_maxiter=None, **kw):
    """The reference's call (rtol, maxiter) on both sides: exit code, iteration count +-1, residual history
    within 1e-10 of ||b|| (north_star).  A map stopped at a loose rtol is only defined to about that rtol, so
    the MAPS are compared at 1e-10 on a second, fully converged solve when ``tight_maxiter`` is given
    (returned instead of the loose ones)."""
    import oracle
    out = []
    for impl, solver in ((oracle, spla.cg), (cm, cm.cg)):
        npix, P, Mbd, A, b, pts = _build(impl, sc, pol, **kw)
        hist = []
        x, info = solver(A, b, M=Mbd, rtol=rtol, maxiter=maxiter,
                         callback=lambda xk: hist.append(np.linalg.norm(b - A * np.asarray(xk))))
        if tight_maxiter:
            x, info_t = solver(A, b, M=Mbd, rtol=TIGHT, maxiter=tight_maxiter)
            assert info_t == 0, "the tight solve did not converge"
        out.append((npix, b, x, info, np.array(hist)))
    (n0, b0, x0, i0, h0), (n1, b1, x1, i1, h1) = out
    assert n0 == n1 and i0 == i1
    assert abs(len(h0) - len(h1)) <= 1
    gc.close(b1, b0, what="rhs")
    k = min(len(h0), len(h1))
    if k:
        assert np.all(np.abs(h1[:k] - h0[:k]) <= 1e-10 * np.linalg.norm(b0)), "residual history differs"
    return x0, x1, h0, h1


def test_config0_ces_standin_unweighted_bd_pcg(cm):
    """configs[0]: one CES, IQU nside=128, A = P^T P, cg(tol=1e-3, maxiter=10) -- the reference's
    src/test_BD_precond_onto_real_data.py:26-47 on the synthetic stand-in for the absent CES file."""
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(600000, nside=128, ndet=4, nx=160, ny=120, samples_per_pixel=12.0, seed=0,
                               flag_turnarounds=True)
    for pol in (1, 3):
        x0, x1, h0, h1 = _solve_both(cm, sc, pol, 1e-3, 10)
        gc.close(x1, x0, rtol=1e-10, what="map pol=%d" % pol)
        assert len(h1) == 1          # exactly preconditioned: one iteration (src/test_BD...:52)


def test_config1_white_noise_bd_pcg(cm):
    """configs[1] at reduced size: white noise blocks, weights fed to M_BD -> 1 iteration."""
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(800000, nside=64, ndet=16, nx=100, ny=60, samples_per_pixel=8.0, seed=2)
    x0, x1, h0, h1 = _solve_both(cm, sc, 3, 1e-10, 20, noise="white")
    gc.close(x1, x0, rtol=1e-10, what="map")
    assert len(h1) == 1
    # mismatched preconditioner (unit-weight M_BD): several iterations, same history on both sides
    x0, x1, h0, h1 = _solve_both(cm, sc, 3, 1e-10, 100, tight_maxiter=300, noise="white", w_from_noise=False)
    gc.close(x1, x0, rtol=1e-10, what="map (unit-weight M_BD)")
    assert len(h1) > 3


@pytest.mark.parametrize("nband", [16, 200])
def test_config2_toeplitz_noise_bd_pcg(cm, nband):
    """configs[2] at reduced size: per-detector banded Toeplitz N^-1 (direct kernel and FFT kernel)."""
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(400000, nside=64, ndet=8, nx=90, ny=50, samples_per_pixel=6.0, seed=4,
                               flag_turnarounds=True)
    bands = synthetic.toeplitz_bands(sc.ndet, nband, seed=1)
    x0, x1, h0, h1 = _solve_both(cm, sc, 3, 1e-9, 200, tight_maxiter=500, noise=bands)
    gc.close(x1, x0, rtol=1e-10, what="map")
    assert len(h1) > 2


def test_config2_subscan_filter_bd_pcg(cm):
    """configs[2]: subscan offset filtering, A = P^T F P (singular: compare A x and histories)."""
    from cosmomap2_b200 import synthetic
    import oracle
    sc = synthetic.raster_scan(400000, nside=64, ndet=8, nx=90, ny=50, samples_per_pixel=6.0, seed=5,
                               flag_turnarounds=True)
    # the reference's own call first (history, count), then a converged solve: A x = b - r, so the two A x
    # differ by at most |r0| + |r1| <= 2e-12 ||b|| plus rounding -- compared at 1e-10
    x0, x1, h0, h1 = _solve_both(cm, sc, 1, 1e-4, 25, filt=True)
    npix, P, Mbd, A, b, pts = _build(oracle, sc, 1, filt=True)
    xs = []
    for impl, solver in ((oracle, spla.cg), (cm, cm.cg)):
        npix_i, P_i, M_i, A_i, b_i, _ = _build(impl, sc, 1, filt=True)
        x, info = solver(A_i, b_i, M=M_i, rtol=1e-12, maxiter=2000)
        assert info == 0
        xs.append(x)
    scale = np.max(np.abs(b))
    assert np.max(np.abs(A * xs[1] - A * xs[0])) <= 1e-10 * scale, "A x differs (null space of P^T F P projected out)"


def test_config3_two_level_from_arnoldi(cm):
    """configs[3] at reduced size: deflation space from the preconditioned Arnoldi (krypy semantics),
    coarse operator by eigendecomposition, M_2lvl PCG; GPU vs oracle on Ritz values, the deflated
    subspace and the iteration counts."""
    import oracle
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(300000, nside=64, ndet=6, nx=60, ny=40, samples_per_pixel=6.0, seed=6,
                               flag_turnarounds=True)
    bands = synthetic.toeplitz_bands(sc.ndet, 24, seed=3, eps=0.9)
    res = []
    for impl, solver in ((oracle, spla.cg), (cm, cm.cg)):
        npix, P, Mbd, A, b, pts = _build(impl, sc, 3, noise=bands)
        n = 3 * npix
        V, H, m = impl.run_krypy_arnoldi(A, np.ones(n), Mbd, 1e-5, maxiter=30, ortho="dmgs")
        theta = np.sort(np.linalg.eigvalsh(H[:H.shape[1], :]))
        r = 6
        thr = 0.5 * (theta[r - 1] + theta[r])
        Z, rr, th = impl.find_ritz_eigenvalues(H, V, threshold=thr, eigenvalues=True)
        assert rr == r
        Az = np.column_stack([A * np.ascontiguousarray(Z[:, i]) for i in range(r)])
        E = impl.CoarseLO(Z, Az, r, apply="eig")
        Zd, AZd = impl.DeflationLO(Z), impl.DeflationLO(Az)
        M2 = Mbd * (impl.lp.IdentityOperator(n) - AZd * E * Zd.T) + Zd * E * Zd.T
        it_bd, it_m2 = [], []
        xb, ib = solver(A, b, M=Mbd, rtol=1e-9, maxiter=300, callback=lambda xk: it_bd.append(1))
        xm, im = solver(A, b, M=M2, rtol=1e-9, maxiter=300, callback=lambda xk: it_m2.append(1))
        assert ib == 0 and im == 0
        xb, ib = solver(A, b, M=Mbd, rtol=TIGHT, maxiter=600)      # converged maps for the 1e-10 comparison
        xm, im = solver(A, b, M=M2, rtol=TIGHT, maxiter=600)
        assert ib == 0 and im == 0
        res.append(dict(theta=theta, Z=np.asarray(Z), xb=xb, xm=xm, nbd=len(it_bd), nm2=len(it_m2)))
    o, g = res
    gc.close(g["theta"][:10], o["theta"][:10], rtol=1e-10, what="Ritz values")
    # same deflation subspace (Z is defined up to rotation/sign): compare projectors on a probe
    probe = np.random.default_rng(0).standard_normal(o["Z"].shape[0])
    po = o["Z"].dot(np.linalg.lstsq(o["Z"], probe, rcond=None)[0])
    pg = g["Z"].dot(np.linalg.lstsq(g["Z"], probe, rcond=None)[0])
    # the 6-dimensional Ritz subspace is defined up to (rounding) / (gap between theta_6 and theta_7, ~1e-2
    # of the spectrum here): 1e-10 holds for the Ritz VALUES above, 1e-9 for the projector
    gc.close(pg, po, rtol=1e-9, what="deflation subspace projector")
    gc.close(g["xb"], o["xb"], rtol=1e-10, what="M_BD solution")
    gc.close(g["xm"], o["xm"], rtol=1e-10, what="M_2lvl solution")
    assert abs(g["nbd"] - o["nbd"]) <= 1 and abs(g["nm2"] - o["nm2"]) <= 1
    assert g["nm2"] <= g["nbd"] + 3     # Ritz vectors after 30 steps are only roughly converged


def test_config0_from_a_ces_file(cm, tmp_path):
    """configs[0] as the reference script runs it (src/test_BD_precond_onto_real_data.py:8-61): a CES
    file in the AnalysisBackend HDF5 schema -> read_from_data(file, pol, npairs=4) -> ProcessTimeSamples
    with the file's obspix -> A = P^T P -> cg(tol=1e-3, maxiter=10) -> reorganize_map -- on the GPU and
    through the oracle, from the same file."""
    import oracle
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(5 * 60000, nside=128, ndet=5, nx=160, ny=120, samples_per_pixel=12.0, seed=4,
                               flag_turnarounds=True)
    obspix = np.unique(sc.pix[sc.pix >= 0])
    idx = np.where(sc.pix >= 0, np.searchsorted(obspix, sc.pix), -1)
    cut = lambda a: [a[b * sc.ns:(b + 1) * sc.ns] for b in range(sc.ndet)]  # noqa: E731
    path = str(tmp_path / "ces.hdf5")
    cm.write_ces_to_hdf5(path, obspix, cut(idx), cut(sc.phi), [np.zeros(sc.ns, dtype=np.int32)] * sc.ndet, sc.ns,
                         sc.sub_len, sc.sub_start, sum_=cut(sc.d), weight_sum=sc.weights, dif=cut(sc.d),
                         weight_dif=sc.weights)
    nside = 128
    for pol in (1, 3):
        res = []
        for impl, solver in ((oracle, spla.cg), (cm, cm.cg)):
            d, weight, phi, pixs, hp_pixs, ground, ces_size = cm.read_from_data(path, pol, npairs=4)
            assert len(d) == 4 * sc.ns and int(ces_size) == sc.ns
            npix = len(hp_pixs)
            pts = impl.ProcessTimeSamples(pixs, npix, obspix=hp_pixs, pol=pol, phi=phi)
            npix, obs = pts.get_new_pixel
            P = impl.SparseLO(npix, len(d), pixs, pol=pol, angle_processed=pts)
            Mbd = impl.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
            A = P.T * P
            b = P.T * d
            x, info = solver(A, b, M=Mbd, rtol=1e-3, maxiter=10)
            hp = (cm if impl is cm else oracle).reorganize_map(x, obs, npix, nside, pol)
            res.append((npix, np.asarray(obs), b, x, info, hp))
        (n0, o0, b0, x0, i0, h0), (n1, o1, b1, x1, i1, h1) = res
        assert n0 == n1 and i0 == 0 and i1 == 0
        gc.exact(o1, o0, "observed HEALPix pixels")
        gc.close(b1, b0, what="rhs from file")
        gc.close(x1, x0, rtol=1e-10, what="map from file, pol=%d" % pol)
        # the HEALPix maps are a permutation of x: same tolerance relative to the same scale (max |x| over
        # all Stokes components, as for x above), and the GPU's own maps are bit-exactly its x
        gc.close(np.concatenate(h1), np.concatenate(h0), rtol=1e-10, what="HEALPix maps")
        for k, a in enumerate(h1):
            assert np.count_nonzero(a) <= n1
            gc.exact(a[np.asarray(o1)], x1[k::pol], "HEALPix map %d is a permutation of x" % k)


def test_device_side_tolerance_and_zero_rhs(cm):
    """cg(A, b, M=M_BD) with x0 = None evaluates SciPy's atol = max(atol, rtol ||b||) and its b = 0 exit in the start
    kernel (no host round trip for ||b||): same exit code, iteration count and iterates as the path that takes the
    norm on the host (x0 = zeros), for rtol-, atol- and maxiter-limited solves; b = 0 returns x = 0, info = 0 like
    scipy.sparse.linalg.cg."""
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(6 * 30000, nside=64, ndet=6, nx=60, ny=30, samples_per_pixel=6.0, seed=3, flag_turnarounds=True)
    npix, P, Mbd, A, b, pts = _build(cm, sc, 3, filt=True)
    n = 3 * npix
    bn = np.linalg.norm(b)
    for kw in (dict(rtol=1e-8, maxiter=500), dict(rtol=0.0, atol=1e-6 * bn, maxiter=500), dict(rtol=1e-30, maxiter=7),
               dict(rtol=1e-4, atol=1e-3 * bn, maxiter=500)):
        r0, r1 = [], []
        x_dev, i_dev = cm.cg(A, b, M=Mbd, residuals=r0, **kw)                      # device-side tolerance
        x_host, i_host = cm.cg(A, b, x0=np.zeros(n), M=Mbd, residuals=r1, **kw)    # norm on the host
        assert i_dev == i_host and len(r0) == len(r1)
        assert np.max(np.abs(np.array(r0) - np.array(r1))) <= 1e-10 * bn
        gc.close(x_dev, x_host, rtol=1e-10, what="device-side vs host-side tolerance %r" % (kw,))
        xs, i_s = spla.cg(A, b, M=Mbd, **kw)
        assert i_s == i_dev
    for M in (Mbd, None):
        x0_, i0_ = cm.cg(A, np.zeros(n), M=M, rtol=1e-8, maxiter=50)
        xs, i_s = spla.cg(A, np.zeros(n), M=M, rtol=1e-8, maxiter=50)
        assert i0_ == 0 and i_s == 0 and not np.any(x0_) and not np.any(xs)
