"""
The reference's own test-suite (/root/reference/tests/*.py), ported to py3/pytest and parametrised
by ``impl`` (the oracle on CPU, the CUDA product on GPU).  Each function cites the test it ports;
inputs come from the seeded port of ``system_setup`` instead of the unseeded original.
Python-2 integer divisions (``nt/nb``) are written ``//``.
"""
import numpy as np
import scipy.linalg as la
import scipy.sparse.linalg as spla


def _setup(impl, nt, npix, nb, seed):
    return impl.system_setup(nt, npix, nb, rng=np.random.default_rng(seed))


def matrix_vector_product(impl):
    """tests/test_matrix_vector_product.py:9-23 -- P^T P 1 == counts."""
    nt, npix = 80, 50
    pairs = impl.pairs_gen(nt, npix, rng=np.random.default_rng(1))
    processd = impl.ProcessTimeSamples(pairs, npix)
    npix = processd.get_new_pixel[0]
    P = impl.SparseLO(npix, nt, pairs)
    x = np.ones(npix)
    y = P.T * (P * x)
    assert np.allclose(y, processd.counts)
    assert np.allclose((P.T * P) * x, processd.counts)


def explicit_blockdiagonal_preconditioner(impl):
    """tests/test_matrix_vector_product.py:26-63 -- M_BD v == inv(block) v per pixel."""
    nt = 10000
    for pol in (3, 1, 2):
        phi = impl.angles_gen(2., nt)
        for i in (10, 50, 100, 300, 600):
            npix = int(i)
            pairs = impl.pairs_gen(nt, npix, rng=np.random.default_rng(1000 * pol + i + 3))
            processd = impl.ProcessTimeSamples(pairs, npix, pol=pol, phi=phi)
            npix = processd.get_new_pixel[0]
            P = impl.SparseLO(npix, nt, pairs, pol=pol, angle_processed=processd)
            x = np.ones(npix * pol)
            v = P.T * (P * x)
            v2 = v * 0.
            Mbd = impl.BlockDiagonalPreconditionerLO(processd, npix, pol=pol)
            if pol == 1:
                v2 = v / processd.counts
            elif pol == 3:
                for j, (s2, c2, cs, s, c, hits) in enumerate(zip(Mbd.sin2, Mbd.cos2, Mbd.sincos, Mbd.sin,
                                                                  Mbd.cos, Mbd.counts)):
                    ainv = la.inv(np.array([[hits, c, s], [c, c2, cs], [s, cs, s2]]))
                    v2[pol * j:pol * j + pol] = np.dot(ainv, v[pol * j:pol * j + pol])
            else:
                for j, (s2, c2, cs) in enumerate(zip(Mbd.sin2, Mbd.cos2, Mbd.sincos)):
                    ainv = la.inv(np.array([[c2, cs], [cs, s2]]))
                    v2[pol * j:pol * j + pol] = np.dot(ainv, v[pol * j:pol * j + pol])
            assert np.allclose(v2, Mbd * v)


def preconditioner_times_matrix_gives_identity(impl):
    """tests/test_matrix_vector_product.py:65-94."""
    nt = 20000
    for pol in (3, 1, 2):
        phi = impl.angles_gen(2., nt)
        for i in (10, 50, 100):
            npix = int(i)
            pairs = impl.pairs_gen(nt, npix, rng=np.random.default_rng(7 * pol + i))
            processd = impl.ProcessTimeSamples(pairs, npix, pol=pol, phi=phi)
            npix = processd.get_new_pixel[0]
            P = impl.SparseLO(npix, nt, pairs, pol=pol, angle_processed=processd)
            Mbd = impl.BlockDiagonalPreconditionerLO(processd, npix, pol=pol)
            x = {1: np.ones(npix), 3: np.tile([0., 0., 1.], npix), 2: np.tile([0., 1.], npix)}[pol]
            v = Mbd * P.T * P * x
            assert np.allclose(v, x)


def block_diagonal_operator(impl):
    """tests/test_block_diagonal_operator.py:8-36."""
    nb = 1
    for nt in (2 ** 14, 2 ** 15):
        for pol in (1, 2, 3):
            for i in (64, 128, 256):
                npix = int(i)
                d, pairs, phi, t, diag = _setup(impl, nt, npix, nb, seed=nt + 10 * pol + i)
                processd = impl.ProcessTimeSamples(pairs, npix, pol=pol, phi=phi)
                npix = processd.get_new_pixel[0]
                P = impl.SparseLO(npix, nt, pairs, pol=pol, angle_processed=processd)
                x = np.ones(pol * npix)
                Mbd = impl.BlockDiagonalPreconditionerLO(processd, npix, pol=pol)
                invMbd = impl.BlockDiagonalLO(processd, npix, pol=pol)
                assert np.allclose(invMbd * x, P.T * P * x)
                assert np.allclose(Mbd * invMbd * x, x)


def spd_properties_block_diagonal_preconditioner(impl):
    """tests/test_block_diagonal_operator.py:39-64 (blocksize list = per-block sizes)."""
    nb = 6
    blocksize = 2 * [500, 400, 124]
    nt = sum(blocksize)
    for pol in (1, 2, 3):
        d, pairs, phi, t, diag = _setup(impl, nt, 64, nb, seed=40 + pol)
        N = impl.BlockLO(blocksize, diag, offdiag=False)
        processd = impl.ProcessTimeSamples(pairs, 64, pol=pol, phi=phi, w=N.diag)
        npix = processd.get_new_pixel[0]
        P = impl.SparseLO(npix, nt, pairs, pol=pol, angle_processed=processd)
        randarray = np.random.default_rng(pol).random(pol * npix)
        A = P.T * N * P
        assert np.allclose(A * randarray, A.T * randarray)
        assert impl.scalprod(randarray, A * randarray) > 0.
        Mbd = impl.BlockDiagonalPreconditionerLO(processd, npix, pol)
        assert np.allclose(Mbd * randarray, Mbd.T * randarray)
        assert impl.scalprod(randarray, Mbd * randarray) > 0.


def toeplitz_vector_products(impl):
    """tests/test_toeplitz_vector_multiplication.py:6-76: associativity for white / Toeplitz N,
    and BlockDiagonalLO == P^T N P for white N with w = N.diag."""
    nb = 6
    blocksize = 2 * [500, 400, 124]
    nt = sum(blocksize)
    for pol in (1, 2, 3):
        for offdiag in (False, True):
            d, pairs, phi, t, diag = _setup(impl, nt, 64, nb, seed=60 + pol)
            N = impl.BlockLO(blocksize, t if offdiag else diag, offdiag=offdiag)
            processd = impl.ProcessTimeSamples(pairs, 64, pol=pol, phi=phi, w=None if offdiag else N.diag)
            npix = processd.get_new_pixel[0]
            P = impl.SparseLO(npix, nt, pairs, pol=pol, angle_processed=processd)
            x = np.ones(pol * npix)
            z = P.T * (N * (P * x))
            z2 = P.T * N * P * x
            assert np.allclose(z2, z)
            if not offdiag:
                PtNP = impl.BlockDiagonalLO(processd, npix, pol=pol)
                assert np.allclose(PtNP * x, z2)


def _deflation_system(impl, nt, npix, nb, pol, seed):
    d, pairs, phi, t, diag = _setup(impl, nt, npix, nb, seed)
    N = impl.BlockLO(nt // nb, t, offdiag=True)
    processd = impl.ProcessTimeSamples(pairs, npix, pol=pol, phi=phi)
    npix = processd.get_new_pixel[0]
    P = impl.SparseLO(npix, nt, pairs, pol=pol, angle_processed=processd)
    Mbd = impl.BlockDiagonalPreconditionerLO(processd, npix, pol=pol)
    B = impl.BlockDiagonalLO(processd, npix, pol=pol)
    A = P.T * N * P
    b = P.T * N * d
    return npix, P, N, Mbd, B, A, b


def deflation_operator(impl):
    """tests/test_deflation_operator.py:6-50."""
    for pol in (1, 2, 3):
        npix, P, N, Mbd, B, A, b = _deflation_system(impl, 1000, 20, 2, pol, seed=80 + pol)
        eigv, Z = spla.eigsh(A, M=B, Minv=Mbd, k=5, which="SM", ncv=50, maxiter=40, tol=1e-4,
                             v0=np.ones(pol * npix))
        r = Z.shape[1]
        assert np.linalg.matrix_rank(Z) == r
        assert la.det(np.asarray(impl.dgemm(Z, Z.T))) != 0
        v = np.ones(r)
        Zd = impl.DeflationLO(Z)
        x = np.ones(pol * npix)
        assert np.allclose(Z.dot(v), Zd * v)
        assert np.allclose(Z.T.dot(x), Zd.H * x)


def coarse_operator(impl):
    """tests/test_coarse_operator.py:6-43 and tests/test_coarse_wclass.py:44-61."""
    for pol in (1, 2, 3):
        npix, P, N, Mbd, B, A, b = _deflation_system(impl, 400 if pol == 1 else 4000, 20, 1, pol, seed=90 + pol)
        eigv, Z = spla.eigsh(A, M=B, Minv=Mbd, k=5, which="SM", ncv=15, tol=1e-5, v0=np.ones(pol * npix))
        r = Z.shape[1]
        Az = Z * 0.
        for i in range(r):
            Az[:, i] = A * Z[:, i]
        invE = impl.CoarseLO(Z, Az, r, apply="eig")
        E = np.asarray(impl.dgemm(Z, Az.T))
        v = np.ones(r)
        y = invE * v
        assert np.allclose(v, np.dot(E, invE * v)) and np.allclose(la.solve(E, v), y)
        evals = la.eigvalsh(invE.to_array())
        assert abs(max(evals) / min(evals)) <= 1.e3


def two_level_preconditioner(impl, cg):
    """tests/test_2level_preconditioner.py:8-53 and tests/test_arnoldi_algorithm.py:51-93:
    M2 A Z_i == Z_i, ||R A Z_i|| <= 1e-10, cg(M2*A, Z_i) exits 0 after exactly 1 iteration."""
    for pol in (1, 2, 3):
        npix, P, N, Mbd, B, A, b = _deflation_system(impl, 500, 40, 1, pol, seed=70 + pol)
        n = pol * npix
        tol = 1e-4
        eigv, Z = spla.eigsh(A, M=B, Minv=Mbd, k=5, v0=np.ones(n), which="SM", ncv=15, tol=1e-10)
        r = Z.shape[1]
        Az = Z * 0.
        for i in range(r):
            Az[:, i] = A * Z[:, i]
        E = impl.CoarseLO(Z, Az, r)
        Zd = impl.DeflationLO(Z)
        I = impl.lp.IdentityOperator(n)
        R = I - A * Zd * E * Zd.T
        M2 = Mbd * R + Zd * E * Zd.T
        for i in range(r):
            assert np.allclose(M2 * A * Z[:, i], Z[:, i])
            assert impl.norm2(R * A * Z[:, i]) <= 1.e-10
            count = []
            x, info = cg(M2 * A, Z[:, i], rtol=tol, maxiter=2, callback=lambda xk: count.append(1))
            assert info == 0
            assert len(count) == 1
