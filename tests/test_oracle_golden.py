"""
Pin the oracle: the NumPy restatement (oracle/operators.py, oracle/krylov.py) must reproduce the
golden vectors that tests/golden/make_golden.py produced by running the reference's own code.
CPU only.
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import oracle
import golden_cases as gc


@pytest.mark.parametrize("pol", [1, 2, 3])
@pytest.mark.parametrize("kind", ["u", "w"])
def test_oracle_pointing_weights_precond(pol, kind):
    gc.check_pointing(oracle, "pointing_pol%d_%s" % (pol, kind))


@pytest.mark.parametrize("pol", [1, 3])
def test_oracle_obspix2_path(pol):
    gc.check_obspix2(oracle, "obspix2_pol%d" % pol)


def test_oracle_noise_ops():
    gc.check_noise_ops(oracle)


def test_oracle_filter_ops():
    gc.check_filter_ops(oracle)


def test_oracle_next_rows_ground_and_legendre_filters():
    gc.check_next_rows(oracle)


def test_oracle_fused_chains():
    """P.T*N*P with short Toeplitz bands, F*P and P.T*F*N*F*P: the oracle against the reference's own run."""
    gc.check_fused_chains(oracle)


def test_oracle_reorganize_map():
    rng = np.random.default_rng(0)
    nside, npix = 8, 40
    obspix = np.sort(rng.choice(12 * nside * nside, npix, replace=False))
    for pol in (1, 2, 3):
        m = rng.standard_normal(pol * npix)
        out = oracle.reorganize_map(m, obspix, npix, nside, pol)
        assert len(out) == pol and all(len(o) == 12 * nside * nside for o in out)
        for k in range(pol):
            assert np.array_equal(out[k][obspix], m[k::pol])
            assert np.count_nonzero(out[k]) == np.count_nonzero(m[k::pol])


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_oracle_solve(pol):
    gc.check_solve(oracle, "solve_pol%d" % pol, spla.cg)


def test_c_loops_match_numpy_restatement():
    """oracle/weave_loops.c (the timed CPU baseline) agrees with oracle/operators.py."""
    from oracle import cloops, operators
    assert cloops.available()
    operators.USE_C_LOOPS = False        # compare the C twin against the pure NumPy restatement
    rng = np.random.default_rng(3)
    nt, npix = 5000, 40
    pix = rng.integers(0, npix, nt)
    pix[rng.random(nt) < 0.05] = -1
    phi = oracle.angles_gen(0.4, nt)
    c, s = np.cos(2 * phi), np.sin(2 * phi)
    w = rng.random(nt) + 0.5
    for pol in (1, 2, 3):
        x = rng.standard_normal(pol * npix)
        d = rng.standard_normal(nt)
        pts = type("A", (), {"cos": c, "sin": s})
        P = oracle.SparseLO(npix, nt, pix, pol=pol, angle_processed=pts)
        assert np.array_equal(cloops.pointing_mult(pix, c, s, pol, x), P * x)
        assert np.array_equal(cloops.pointing_rmult(pix, c, s, pol, d, npix), P.T * d)
    counts, cosine, sine, cos2, sin2, sincos = cloops.moments(pix, w, c, s, 3, npix)
    pp = pix.copy()
    pts = oracle.ProcessTimeSamples(pp, npix, pol=3, phi=phi, w=w)
    keep = np.asarray(pts.old2new) >= 0
    for a, b in ((counts, pts.counts), (cosine, pts.cosine), (sine, pts.sine), (cos2, pts.cos2),
                 (sin2, pts.sin2), (sincos, pts.sincos)):
        assert np.array_equal(a[keep], b)
    Mbd = oracle.BlockDiagonalPreconditionerLO(pts, pts.get_new_pixel[0], pol=3)
    x = rng.standard_normal(3 * pts.get_new_pixel[0])
    y = cloops.bd_apply(pts.get_new_pixel[0], 3, pts.counts, pts.cosine, pts.sine, pts.cos2,
                        pts.sin2, pts.sincos, x)
    gc.close(y, Mbd * x, rtol=1e-13, what="bd apply")
    operators.USE_C_LOOPS = True


@pytest.mark.parametrize("use_c", [False, True])
def test_oracle_golden_both_restatements(use_c):
    """Both restatements (NumPy and C) reproduce the reference's golden vectors."""
    from oracle import operators
    old = operators.USE_C_LOOPS
    operators.USE_C_LOOPS = use_c
    try:
        gc.check_pointing(oracle, "pointing_pol3_w")
        gc.check_pointing(oracle, "pointing_pol2_u")
        gc.check_pointing(oracle, "pointing_pol1_w")
    finally:
        operators.USE_C_LOOPS = old
