"""The dense tall-skinny contractions of the deflation path (csrc/dense_z.cu, fp64 tensor cores) against
NumPy: E = Z^T (A Z) in one pass (reference: dgemm(Z, Az.T), interfaces/linearoperators.py:1019) and the
Ritz-vector assembly Z = V U (interfaces/deflationlib.py:204-219), at ragged sizes and alignments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dense():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cosmomap2_b200 import dense
    return dense


@pytest.mark.parametrize("n,r1,r2", [(100000, 32, 32), (100003, 5, 7), (17, 3, 3), (1, 1, 1), (4096, 8, 32), (33333, 17, 9),
                                     (20001, 40, 33), (15, 64, 64), (250000, 24, 24)])
def test_gram_equals_numpy(dense, n, r1, r2):
    from cosmomap2_b200 import _device as dv
    rng = np.random.default_rng(n + r1 + r2)
    X, Y = rng.standard_normal((r1, n)), rng.standard_normal((r2, n))
    got = dv.to_host(dense.gram(dv.to_dev_f64(X), dv.to_dev_f64(Y)))
    ref = X.dot(Y.T)
    scale = np.sqrt(n) * 3
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 1e-13 * scale * max(1.0, np.max(np.abs(ref)) / scale)
    # deterministic: the same bits on a second call
    again = dv.to_host(dense.gram(dv.to_dev_f64(X), dv.to_dev_f64(Y)))
    assert np.array_equal(got, again)


def test_gram_unaligned_views(dense):
    """Row views with an odd leading dimension / offset take the scalar-load path."""
    import torch
    from cosmomap2_b200 import _device as dv
    rng = np.random.default_rng(0)
    n = 5001
    base = dv.to_dev_f64(rng.standard_normal((6, n + 3)))
    X = base[:, 1:n + 1]                       # ld = n + 3 (odd), 8-byte offset
    got = dv.to_host(dense.gram(X, X))
    Xh = dv.to_host(base)[:, 1:n + 1]
    assert np.max(np.abs(got - Xh.dot(Xh.T))) <= 1e-11


@pytest.mark.parametrize("n,m,r", [(100000, 50, 32), (100001, 7, 5), (31, 3, 2), (1, 1, 1), (4097, 33, 17), (20000, 120, 40),
                                   (65536, 300, 32)])
def test_combine_equals_numpy(dense, n, m, r):
    from cosmomap2_b200 import _device as dv
    rng = np.random.default_rng(n + m + r)
    V, U = rng.standard_normal((m, n)), rng.standard_normal((m, r))
    got = dv.to_host(dense.combine(dv.to_dev_f64(V), U))
    ref = U.T.dot(V)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 1e-13 * np.sqrt(m) * 10 * max(1.0, np.max(np.abs(ref)))


def test_coarse_operator_one_pass_equals_oracle(dense):
    """CoarseLO's E through the one-pass kernel against the oracle's dgemm on the golden deflation space."""
    import cosmomap2_b200 as cm
    import oracle
    import golden_cases as gc
    g = gc.load("solve_pol3")
    Z, Az = g["Z"], g["Az"]
    r = Z.shape[1]
    E = cm.CoarseLO(Z, Az, r)
    gc.close(E.E, g["E"], rtol=1e-12, what="E = Z^T A Z")
    gc.close(cm.dgemm(Z, Az.T), oracle.dgemm(Z, Az.T), rtol=1e-12, what="dgemm")


def test_scan_coarse_space_deflates_the_offset_filter_modes(dense):
    """The a-priori scan-aligned coarse space: with r = number of map rows it captures the near-null space of
    M_BD P^T F P (maps constant along every subscan) and the two-level PCG needs a fraction of M_BD's
    iterations (CPU prototype of the same recipe on a 24-row scan: 42 -> 10); the GPU solve equals the oracle's
    with the same Z."""
    import scipy.sparse.linalg as spla
    import cosmomap2_b200 as cm
    import oracle
    from cosmomap2_b200 import synthetic, _device as dv
    sc = synthetic.raster_scan(240000, nside=64, ndet=8, nx=48, ny=24, samples_per_pixel=8.0, seed=1, flag_turnarounds=True)
    pol, r = 3, 24
    out = {}
    Zt_host = None
    for label, impl, solver in (("gpu", cm, cm.cg), ("oracle", oracle, spla.cg)):
        pix = sc.pix.astype(np.int64)
        pts = impl.ProcessTimeSamples(pix, sc.npix_full, obspix=np.arange(sc.npix_full), pol=pol, phi=sc.phi)
        npix = pts.get_new_pixel[0]
        n = pol * npix
        P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix)
        Mbd = impl.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
        A = P.T * F * P
        rng = np.random.default_rng(0)
        d = P * rng.standard_normal(n) + 0.5 * rng.standard_normal(sc.nt)
        b = P.T * (F * d)
        if label == "gpu":
            Zt = cm.scan_coarse_space(P, r, sc.ns, A=A, Mbd=Mbd, smooth=2)
            Zt_host = dv.to_host(Zt)
            # every observed pixel belongs to exactly one band, intensity component only
            assert np.array_equal(Zt_host.sum(axis=0)[0::3], np.ones(npix))
            assert np.count_nonzero(Zt_host[:, 1::3]) == 0 and np.count_nonzero(Zt_host[:, 2::3]) == 0
        Z = Zt_host.T.copy()
        AZ = np.column_stack([A * Z[:, i] for i in range(r)])
        E = impl.CoarseLO(Z, AZ, r, apply="eig")
        Zd, AZd = impl.DeflationLO(Z), impl.DeflationLO(AZ)
        M2 = Mbd * (impl.lp.IdentityOperator(n) - AZd * E * Zd.T) + Zd * E * Zd.T
        it_bd, it_m2 = [], []
        xb, ib = solver(A, b, M=Mbd, rtol=1e-8, maxiter=500, callback=lambda xk: it_bd.append(1))
        xm, im = solver(A, b, M=M2, rtol=1e-8, maxiter=500, callback=lambda xk: it_m2.append(1))
        assert ib == 0 and im == 0
        out[label] = (len(it_bd), len(it_m2), A * xm if label == "oracle" else None, xm)
    (gb, gm, _, xg), (ob, om, Axo, xo) = out["gpu"], out["oracle"]
    assert abs(gb - ob) <= 1 and abs(gm - om) <= 1
    assert gm <= 0.4 * gb, "the scan-aligned coarse space must cut the iteration count (%d vs %d)" % (gm, gb)


@pytest.mark.parametrize("pol", [1, 3])
def test_banded_two_level_apply_equals_dense(dense, pol):
    """The two-level apply in banded form (cm2_m2_banded_apply: band index + three entries of AZ per map element)
    for the scan coarse space against the dense apply (cm2_m2_apply) on the same Z, AZ, E; a generic (dense) Z and
    an A Z that leaves the three bands are refused by the exact check and keep the dense apply."""
    import torch
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import synthetic, linearoperators as lo, _device as dv
    sc = synthetic.raster_scan(8 * 45000, nside=64, ndet=8, nx=60, ny=96, samples_per_pixel=8.0, seed=2, flag_turnarounds=True)
    r = 12
    pix = sc.pix.astype(np.int64)
    pts = cm.ProcessTimeSamples(pix, sc.npix_full, obspix=np.arange(sc.npix_full), pol=pol, phi=sc.phi)
    npix = pts.get_new_pixel[0]
    n = pol * npix
    P = cm.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
    F = cm.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    A = P.T * F * P
    Zt = cm.scan_coarse_space(P, r, sc.ns, A=A, Mbd=Mbd, smooth=2)
    AZt = torch.stack([A._apply(Zt[i]) for i in range(r)])
    # A Z by probing (4 applies to sums of every 4th column instead of 12) equals the column-by-column products; a
    # space whose bands are narrower than the reach of A (96 bands on 96 rows) fails the zero check and falls back
    n0 = A.nMatvec
    AZp = cm.coarse_products(A, Zt, pol)
    assert A.nMatvec - n0 == 5, "4 probes + 1 checksum apply expected, not the column-by-column fallback"
    assert float((AZp - AZt).abs().max() / AZt.abs().max()) < 1e-13
    Zt_thin = cm.scan_coarse_space(P, 96, sc.ns, A=A, Mbd=Mbd, smooth=2)
    AZ_thin = cm.coarse_products(A, Zt_thin, pol)
    for k in (0, 17, 95):
        ref_k = A._apply(Zt_thin[k])
        assert float((AZ_thin[k] - ref_k).abs().max()) <= 1e-13 * max(float(ref_k.abs().max()), 1e-300)
    E = cm.CoarseLO(Zt.t(), AZt.t(), r, apply="eig")
    Zd, AZd = cm.DeflationLO(Zt.t()), cm.DeflationLO(AZt.t())
    v = dv.to_dev_f64(np.random.default_rng(3).standard_normal(n))
    out = {}
    for banded in (True, False):
        lo.M2_BANDED = banded
        try:
            M2 = Mbd * (cm.lp.IdentityOperator(n) - AZd * E * Zd.T) + Zd * E * Zd.T
            out[banded] = M2._apply(v).clone()
        finally:
            lo.M2_BANDED = True
    T = lo.TwoLevelPreconditionerLO(Mbd, Zd, AZd, E)
    assert T._banded is not None, "the scan coarse space must be recognised as banded"
    band = dv.to_host(T._banded[0])
    assert band.min() >= 0 and band.max() == r - 1
    err = float((out[True] - out[False]).abs().max() / out[False].abs().max())
    assert err < 1e-13, "banded vs dense two-level apply: %.2e" % err
    assert torch.equal(T._apply(v), T._apply(v)), "the banded apply is deterministic"
    # refused: a dense Z; an AZ with an entry outside the three bands
    Zr = cm.DeflationLO(torch.randn((r, n), dtype=torch.float64, device="cuda").t())
    assert lo.TwoLevelPreconditionerLO(Mbd, Zr, AZd, E)._banded is None
    AZbad = AZt.clone()
    far = int(np.nonzero(band == 0)[0][0])
    AZbad[r // 2, pol * far] = 1.0
    assert lo.TwoLevelPreconditionerLO(Mbd, Zd, cm.DeflationLO(AZbad.t()), E)._banded is None


@pytest.mark.parametrize("pol", [1, 3])
def test_coarse_products_refuses_a_coupling_that_skips_the_checked_bands(pol):
    """A couples band b to b-3 and b+3 only (nothing at distance 1 and 2): with 4 colours the zero check of the
    colour of b+2 passes, and the entry of band b-3 would be booked on band b+1 (same colour).  The weighted checksum
    must notice and coarse_products must return the column-by-column products; a nearest-band coupling on the same
    space goes through the probing path (counted by the number of A applies)."""
    import torch
    import cosmomap2_b200 as cm
    from cosmomap2_b200 import linop as lp
    r, w = 12, 50
    npix = r * w
    n = pol * npix
    Zt = torch.zeros((r, n), dtype=torch.float64, device="cuda")
    for k in range(r):
        Zt[k, pol * k * w:pol * (k + 1) * w:pol] = 1.0
    g = torch.Generator(device="cuda").manual_seed(5)
    dg = 1.0 + torch.rand(n, dtype=torch.float64, device="cuda", generator=g)

    def coupled(shift):
        def mv(x):
            return dg * x + 0.25 * (torch.roll(x, pol * shift * w) + torch.roll(x, -pol * shift * w))
        return lp.LinearOperator(n, n, matvec=mv, symmetric=True, device=True)

    for shift, applies in ((1, 5), (3, 5 + r)):
        A = coupled(shift)
        n0 = A.nMatvec
        AZ = cm.coarse_products(A, Zt, pol)
        assert A.nMatvec - n0 == applies, (shift, A.nMatvec - n0)
        for k in range(r):
            assert torch.equal(AZ[k], A._apply(Zt[k])), (shift, k)
