"""GPU parity against the oracle on seeded inputs at sizes the oracle finishes in seconds, plus
size-independent properties (hit counts, symmetry, M_BD A = I for white noise, fused == unfused,
atomic == sorted P^T) and the edge cases the reference's tests exercise."""
import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import cosmomap2_b200
    return cosmomap2_b200


def _raster(nt=400000, ndet=8, seed=3, **kw):
    from cosmomap2_b200 import synthetic
    return synthetic.raster_scan(nt, nside=64, ndet=ndet, nx=90, ny=50, samples_per_pixel=6.0, seed=seed, **kw)


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_raster_scan_against_oracle(cm, pol):
    import oracle
    sc = _raster(flag_turnarounds=True)
    res = []
    for impl in (oracle, cm):
        pix = sc.pix.astype(np.int64)
        N = impl.BlockLO(sc.ns, sc.weights, offdiag=False)
        pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        Mbd = impl.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
        rng = np.random.default_rng(5)
        x = rng.standard_normal(pol * npix)
        A = P.T * N * P
        res.append(dict(pix=pix, npix=npix, obspix=np.asarray(pts.get_new_pixel[1]), Px=P * x,
                        Ax=A * x, b=P.T * (N * sc.d), MAx=Mbd * (A * x), x=x))
    o, g = res
    assert o["npix"] == g["npix"]
    gc.exact(g["pix"], o["pix"], "relabelled pixels")
    gc.exact(g["obspix"], o["obspix"], "obspix")
    for k in ("Px", "Ax", "b", "MAx"):
        gc.close(g[k], o[k], what=k)
    # white noise with w = N.diag: M_BD A = I on every kept pixel (tests/test_matrix_vector_product.py:65-94)
    gc.close(g["MAx"], g["x"], rtol=1e-9, what="M_BD A x = x")


def test_hit_counts_bit_exact(cm):
    import oracle
    sc = _raster(nt=300000)
    pix_o = sc.pix.astype(np.int64)
    pix_g = sc.pix.astype(np.int64)
    po = oracle.ProcessTimeSamples(pix_o, sc.npix_full)
    pg = cm.ProcessTimeSamples(pix_g, sc.npix_full)
    n = pg.get_new_pixel[0]
    assert n == po.get_new_pixel[0]
    ref = np.bincount(pix_o[pix_o >= 0], minlength=n)
    assert np.array_equal(pg.hits(), ref)
    assert np.array_equal(np.asarray(pg.counts).astype(np.int64), ref)
    # P^T P 1 == counts (tests/test_matrix_vector_product.py:9-23)
    P = cm.SparseLO(n, sc.nt, pix_g)
    y = P.T * (P * np.ones(n))
    assert np.array_equal(y, ref.astype(np.float64))
    assert np.array_equal(P.hits(), ref)


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_fused_equals_unfused_and_sorted(cm, pol):
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(nt=250000, ndet=5, flag_turnarounds=True)
    pix = sc.pix.astype(np.int64)
    N = cm.BlockLO(sc.ns, sc.weights)
    pts = cm.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
    F = cm.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix)
    x = np.random.default_rng(1).standard_normal(pol * npix)
    for mid in (N, F, None):
        A = (P.T * mid * P) if mid is not None else (P.T * P)
        fused = A * x
        assert any(isinstance(f, (lo._FusedWhiteA, lo._FusedFilterA)) for f in A.planned())
        unfused = P.T * ((mid * (P * x)) if mid is not None else (P * x))
        gc.close(fused, unfused, what="fused vs unfused")
        if mid is not F:                              # FilterLO has no rmatvec (linearoperators.py:277-278)
            gc.close(A.T * x, fused, what="A symmetric")
        else:                                         # P^T F P is symmetric all the same
            y = np.random.default_rng(2).standard_normal(pol * npix)
            assert abs(y.dot(A * x) - x.dot(A * y)) <= 1e-9 * abs(y.dot(A * x)) + 1e-9
    d = sc.d
    gc.close(P.rmult_sorted(d).cpu().numpy(), P.T * d, what="sorted P^T vs atomic P^T")
    # the sorted path is deterministic: two runs are bit-identical
    assert np.array_equal(P.rmult_sorted(d).cpu().numpy(), P.rmult_sorted(d).cpu().numpy())


def test_edge_cases(cm):
    # empty TOD
    pts = cm.ProcessTimeSamples(np.zeros(0, dtype=np.int64), 5)
    assert pts.get_new_pixel[0] == 0
    # everything flagged
    pix = np.full(1000, -1, dtype=np.int64)
    pts = cm.ProcessTimeSamples(pix, 7, pol=3, phi=np.zeros(1000))
    assert pts.get_new_pixel[0] == 0 and np.all(pix == -1)
    # ragged sizes around the 256-sample tile and 8-sample chunk
    import oracle
    for nt in (1, 7, 8, 9, 255, 256, 257, 1023):
        rng = np.random.default_rng(nt)
        pix = rng.integers(0, 5, nt)
        phi = rng.random(nt)
        po, pg = pix.copy(), pix.copy()
        a = oracle.ProcessTimeSamples(po, 5, pol=3, phi=phi)
        b = cm.ProcessTimeSamples(pg, 5, pol=3, phi=phi)
        assert a.get_new_pixel[0] == b.get_new_pixel[0]
        gc.exact(pg, po)
        n = a.get_new_pixel[0]
        if n == 0:
            continue
        Po = oracle.SparseLO(n, nt, po, pol=3, angle_processed=a)
        Pg = cm.SparseLO(n, nt, pg, pol=3, angle_processed=b)
        x = rng.standard_normal(3 * n)
        d = rng.standard_normal(nt)
        gc.close(Pg * x, Po * x)
        gc.close(Pg.T * d, Po.T * d)
        gc.close((Pg.T * Pg) * x, Po.T * (Po * x))
    # errors the reference raises
    with pytest.raises(RuntimeError):
        cm.SparseLO(3, 3, np.arange(3), pol=4)
    P = cm.SparseLO(3, 3, np.arange(3))
    with pytest.raises(cm.lp.ShapeError):
        P * np.ones(5)
    N = cm.BlockLO(4, [1.0, 2.0])
    with pytest.raises(cm.lp.ShapeError):
        N * np.ones(9)


def test_toeplitz_wide_band_and_variable_blocks(cm):
    import oracle
    rng = np.random.default_rng(9)
    sizes = 2 * [500, 400, 124]                       # tests/test_toeplitz_vector_multiplication.py:11
    nt = sum(sizes)
    v = rng.standard_normal(nt)
    t = [rng.random(3) for _ in sizes]
    gc.close(cm.BlockLO(sizes, t, offdiag=True) * v, oracle.BlockLO(sizes, t, offdiag=True) * v)
    tw = rng.random(len(sizes))
    gc.close(cm.BlockLO(sizes, tw) * v, oracle.BlockLO(sizes, tw) * v)
    # band wider than the tile and wider than a block
    n = 5000
    v = rng.standard_normal(n)
    for L in (1500, 4096):
        a = rng.random(L) / L
        a[0] = 1.0
        gc.close(cm.ToeplitzLO(a, n) * v, oracle.ToeplitzLO(a, n) * v, what="Toeplitz L=%d" % L)
    a = rng.random(300)
    gc.close(cm.ToeplitzLO(a, 100) * v[:100], oracle.ToeplitzLO(a, 100) * v[:100], what="band > block")


def test_krypy_style_arnoldi_and_two_level(cm):
    """Preconditioned Arnoldi (krypy semantics): V^T P = I, M A V_m = V_{m+1} H, Ritz values against
    dense eigh of the explicit pencil; then M_2lvl built from it deflates the small modes."""
    import scipy.linalg as la
    g = gc.load("solve_pol3")
    pol, npix, P, N, Mbd, B, A, b = gc.build_solve_system(cm, g)
    n = pol * npix
    V, H, m = cm.run_krypy_arnoldi(A, np.ones(n), Mbd, 1e-5, maxiter=40)
    Vg, Hg, Pg = cm.krypy_arnoldi(A, np.ones(n), M=Mbd, maxiter=40)
    assert np.max(np.abs(Vg.T.dot(Pg) - np.eye(Vg.shape[1]))) < 1e-10
    k = Hg.shape[1]
    MAV = np.column_stack([Mbd * (A * Vg[:, j]) for j in range(k)])
    gc.close(MAV, Vg.dot(Hg), rtol=1e-9, what="M A V_m = V_{m+1} H")
    theta = la.eigvalsh(Hg[:k, :k])
    Ad, Bd = A.to_array(), B.to_array()
    lam = la.eigh(Ad, Bd, eigvals_only=True)
    assert abs(theta.min() - lam.min()) < 1e-3 * abs(lam.min()) + 1e-8
    assert theta.max() <= lam.max() * (1 + 1e-8)
    Z, r, th = cm.find_ritz_eigenvalues(Hg, Vg, threshold=np.sort(theta)[4] * 1.0001, eigenvalues=True)
    assert r == 5
    Az = np.column_stack([A * Z[:, i] for i in range(r)])
    E = cm.CoarseLO(Z, Az, r, apply="eig")
    Zd, AZd = cm.DeflationLO(Z), cm.DeflationLO(Az)
    M2 = Mbd * (cm.lp.IdentityOperator(n) - AZd * E * Zd.T) + Zd * E * Zd.T
    from cosmomap2_b200.linearoperators import TwoLevelPreconditionerLO
    M2f = TwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
    v = np.random.default_rng(0).standard_normal(n)
    gc.close(M2 * v, M2f * v, what="algebraic M2 == fused M2")
    for i in range(r):
        assert np.allclose(M2 * Az[:, i], Z[:, i])
    it_bd, it_m2 = [], []
    cm.cg(A, b, M=Mbd, rtol=1e-10, maxiter=500, residuals=it_bd)
    cm.cg(A, b, M=M2, rtol=1e-10, maxiter=500, residuals=it_m2)
    assert len(it_m2) <= len(it_bd)


@pytest.mark.parametrize("pair_min", [None, 2])
def test_toeplitz_fft_path_equals_direct_path(cm, pair_min):
    """Overlap-save FFT kernel vs the direct shared-memory kernel vs the oracle, multi-block, ragged blocks.
    ``pair_min = 2`` forces every band through the 32768-sample windows on 2-CTA clusters (by default bands of 3000
    coefficients and more take them: the L = 3000, 4096 and 6000 cases below)."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    rng = np.random.default_rng(21)
    for sizes, L in ((3 * [20000], 200), ([30000, 9000, 41000], 1000), ([50000], 4096), ([70001, 33000, 5], 3000),
                     ([60000, 21001], 6000)):
        nt = sum(sizes)
        v = rng.standard_normal(nt)
        t = [np.concatenate([[1.0 + rng.random()], -0.3 * rng.random(L - 1) / L]) for _ in sizes]
        old = lo.TOEPLITZ_FFT_MIN_BAND, lo.TOEPLITZ_FFT_PAIR_MIN_BAND
        try:
            lo.TOEPLITZ_FFT_MIN_BAND = 10 ** 9
            y_direct = cm.BlockLO(sizes, t, offdiag=True) * v
            lo.TOEPLITZ_FFT_MIN_BAND = 2
            if pair_min is not None:
                lo.TOEPLITZ_FFT_PAIR_MIN_BAND = pair_min
            N = cm.BlockLO(sizes, t, offdiag=True)
            y_fft = N * v
            assert N._fft is not None and N._fft.ok
            assert N._fft.pair == (1 if L >= lo.TOEPLITZ_FFT_PAIR_MIN_BAND else 0)
            y_fft2 = N * v                                         # second call: tables already built
        finally:
            lo.TOEPLITZ_FFT_MIN_BAND, lo.TOEPLITZ_FFT_PAIR_MIN_BAND = old
        y_ref = oracle.BlockLO(sizes, t, offdiag=True) * v
        gc.close(y_direct, y_ref, what="direct Toeplitz L=%d" % L)
        gc.close(y_fft, y_ref, what="FFT Toeplitz L=%d" % L)
        gc.exact(y_fft2, y_fft, "FFT Toeplitz is deterministic")


def test_mask_differs_only_on_rounding_knife_edge_pixels(cm):
    """The reference's good-pixel rule evaluates sqrt(tr^2/4 - det) (process_ces.py:544-549).  For a
    pixel whose QU block is isotropic to rounding (hit over whole HWP periods) that argument is
    +-1e-17: NaN (pixel dropped) or not, depending on the summation order of the moments.  The
    GPU mask may differ from the serial oracle ONLY on such pixels."""
    import oracle
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(600000, nside=128, ndet=4, nx=160, ny=120, samples_per_pixel=12.0, seed=0,
                               flag_turnarounds=True, hwp_jitter=0.0)
    po = oracle.ProcessTimeSamples(sc.pix.astype(np.int64), sc.npix_full, pol=3, phi=sc.phi)
    pg = cm.ProcessTimeSamples(sc.pix.astype(np.int64), sc.npix_full, pol=3, phi=sc.phi)
    diff = np.setxor1d(np.asarray(po.mask), np.asarray(pg.mask))
    pix = sc.pix.astype(np.int64)
    g = pix >= 0
    c, s = np.cos(2 * sc.phi[g]), np.sin(2 * sc.phi[g])
    n = sc.npix_full
    c2 = np.bincount(pix[g], weights=c * c, minlength=n)
    s2 = np.bincount(pix[g], weights=s * s, minlength=n)
    cs = np.bincount(pix[g], weights=s * c, minlength=n)
    tr = c2 + s2
    arg = tr * tr / 4 - (c2 * s2 - cs * cs)
    knife = np.abs(arg[diff]) <= 1e-12 * tr[diff] ** 2
    assert knife.all(), "mask differs on %d well-determined pixels" % int((~knife).sum())
    assert len(diff) < 0.01 * len(po.mask)


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_filter_run_table_path_equals_two_pass_path(cm, pol):
    """P^T F P through the run-compressed table (one TOD pass) == the two-pass kernel == oracle."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(nt=300000, ndet=6, seed=8, flag_turnarounds=True)
    rng = np.random.default_rng(3)
    flagged = rng.random(sc.nt) < 0.02                 # flags inside subscans too
    sc.pix[flagged] = -1
    res = {}
    for name, impl, table in (("oracle", oracle, None), ("table", cm, True), ("twopass", cm, False)):
        pix = sc.pix.astype(np.int64)
        pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix)
        x = np.random.default_rng(4).standard_normal(pol * npix)
        if table is None:
            res[name] = P.T * (F * (P * x))
            continue
        old = lo.FILTER_RUN_TABLE
        lo.FILTER_RUN_TABLE = table
        try:
            A = P.T * F * P
            res[name] = A * x
            fused = [f for f in A.planned() if isinstance(f, lo._FusedFilterA)][0]
            assert bool(fused._runs) == table
        finally:
            lo.FILTER_RUN_TABLE = old
    gc.close(res["table"], res["oracle"], what="run-table path")
    gc.close(res["twopass"], res["oracle"], what="two-pass path")


@pytest.mark.parametrize("pol,weighted", [(1, False), (2, True), (3, False), (3, True)])
def test_strict_parity_setup_is_bit_identical(cm, pol, weighted):
    """STRICT_PARITY: the six moment arrays, cos/sin, the mask, old2new, obspix and the relabelled
    pixels are BIT-identical to the serial reference order -- even on the exactly periodic scan
    whose isotropic pixels sit on the rounding knife edge of the reference's mask."""
    import oracle
    from oracle import operators
    from cosmomap2_b200 import synthetic, process_ces
    sc = synthetic.raster_scan(600000, nside=128, ndet=4, nx=160, ny=120, samples_per_pixel=12.0, seed=0,
                               flag_turnarounds=True, hwp_jitter=0.0)
    old_c, old_s = operators.USE_C_LOOPS, process_ces.STRICT_PARITY
    operators.USE_C_LOOPS = False
    process_ces.STRICT_PARITY = True
    try:
        res = []
        for impl in (oracle, cm):
            pix = sc.pix.astype(np.int64)
            N = impl.BlockLO(sc.ns, sc.weights)
            pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag if weighted else None)
            res.append((pix, pts))
    finally:
        operators.USE_C_LOOPS, process_ces.STRICT_PARITY = old_c, old_s
    (po, a), (pg, b) = res
    assert a.get_new_pixel[0] == b.get_new_pixel[0]
    assert np.array_equal(pg, po)
    assert np.array_equal(np.asarray(a.mask), np.asarray(b.mask))
    assert np.array_equal(np.asarray(a.old2new), np.asarray(b.old2new))
    assert np.array_equal(np.asarray(a.get_new_pixel[1]), np.asarray(b.get_new_pixel[1]))
    names = {1: ["counts"], 2: ["cos2", "sin2", "sincos", "cos", "sin"],
             3: ["counts", "cosine", "sine", "cos2", "sin2", "sincos", "cos", "sin"]}[pol]
    for nm in names:
        assert np.array_equal(np.asarray(getattr(a, nm)), np.asarray(getattr(b, nm))), nm


def test_pcg_cooperative_tail_equals_three_kernel_tail(cm):
    """cm2_pcg_bd_iter (one cooperative launch) and the three-kernel fallback give the same iterates."""
    import torch
    from cosmomap2_b200.pcg import PCG
    from cosmomap2_b200 import _device as dv
    g = gc.load("solve_pol3")
    pol, npix, P, N, Mbd, B, A, b = gc.build_solve_system(cm, g)
    n = pol * npix
    bd = dv.to_dev_f64(b)
    outs = []
    for coop in (True, False):
        s = PCG(A, Mbd, n)
        s._coop = coop
        s.start(bd, None, 0.0)
        hist = []
        for _ in range(12):
            s.step_async()
            hist.append(s.state()[0])
        assert s._coop == coop
        outs.append((dv.to_host(s.x), np.array(hist)))
    gc.close(outs[0][0], outs[1][0], rtol=1e-12, what="x after 12 iterations")
    bn = np.linalg.norm(b)
    assert np.max(np.abs(outs[0][1] - outs[1][1])) <= 1e-12 * bn, "residual history (relative to ||b||)"
    k = min(6, len(g["cg_hist"]))
    # the recurrence's ||r|| against the TRUE residual norms of the reference's SciPy run, relative to ||b||
    assert np.max(np.abs(outs[0][1][:k] - g["cg_hist"][:k])) <= 1e-10 * bn, "history vs the reference's SciPy run"
