"""GPU parity for the SURVEY section 8(f) rows: the Legendre path of FilterLO, GroundFilterLO, the
fused P^T F_K P A-matvec and reorganize_map -- against the fixture generated from the reference's
own code (tests/golden/next_rows.npz) and against the oracle on seeded inputs with the edge cases
of the domain (flagged, fully flagged and nearly fully flagged subscans, gaps, several CES,
unsorted subscan tables, subscans longer than the shared-memory window).  fp64 bar: 1e-10."""
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


@pytest.fixture(params=["tma", "ldg"])
def variant(request, cm):
    """Both variants of the subscan kernel: (B) subscans streamed through shared memory by the TMA
    engine (the default when they fit) and (A) staged by ordinary loads."""
    from cosmomap2_b200 import _cabi
    old = _cabi.call("cm2_filter_poly_set_tma", 1 if request.param == "tma" else 0)
    yield request.param
    _cabi.call("cm2_filter_poly_set_tma", old)


def test_golden_ground_and_legendre_filters(cm, variant):
    gc.check_next_rows(cm)


def _subscan_table(rng, ns, nsub, gap=True):
    """nsub subscans covering [0, ns) with random lengths and (optionally) gaps between them."""
    cuts = np.sort(rng.choice(np.arange(1, ns), size=2 * nsub - 1, replace=False))
    edges = np.concatenate([[0], cuts, [ns]])
    starts, lens = [], []
    for k in range(nsub):
        a, b = edges[2 * k], edges[2 * k + 1]
        if not gap:
            b = edges[2 * k + 2] if 2 * k + 2 < len(edges) else ns
        starts.append(a)
        lens.append(b - a)
    return np.array(lens, dtype=np.int64), np.array(starts, dtype=np.int64)


def _flag_patterns(rng, pix, starts, lens, order):
    """Exercise every branch of polyfilter (:183-201) on the first detector's subscans."""
    k = 0
    a, n = starts[k], lens[k]
    pix[a:a + n] = -1                                     # fully flagged
    k = 1
    a, n = starts[k], lens[k]
    pix[a:a + n] = -1
    keep = rng.choice(n, size=min(order, n), replace=False)
    pix[a + keep] = 1                                     # exactly `order` unflagged samples -> skipped
    k = 2
    a, n = starts[k], lens[k]
    pix[a:a + n] = -1
    keep = rng.choice(n, size=min(order + 1, n), replace=False)
    pix[a + keep] = 2                                     # order+1 unflagged samples -> interpolated exactly
    k = 3
    a, n = starts[k], lens[k]
    pix[a:a + n] = np.abs(pix[a:a + n])                   # no flag at all -> the literal non-orthogonal sum
    k = 4
    a, n = starts[k], lens[k]
    pix[a:a + n] = np.abs(pix[a:a + n])
    pix[a:a + n // 2] = -1                                # first half flagged: unflagged samples cluster


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5, 6, 7])
def test_subscan_filter_against_oracle(cm, order, variant):
    import oracle
    rng = np.random.default_rng(100 + order)
    nsamples, nbolos = [6000, 4500], [3, 2]
    subs, tst = [], []
    for ns in nsamples:
        L, S = _subscan_table(rng, ns, 9)
        subs.append(L)
        tst.append(S)
    nt = sum(a * b for a, b in zip(nsamples, nbolos))
    pix = rng.integers(0, 50, size=nt).astype(np.int64)
    pix[rng.random(nt) < 0.1] = -1
    _flag_patterns(rng, pix, tst[0], subs[0], max(order, 1))
    d = rng.standard_normal(nt) + 3.0 * np.sin(np.arange(nt) / 700.0) + 10.0
    res = []
    for impl in (oracle, cm):
        F = impl.FilterLO(nt, [subs, tst], nsamples, nbolos, pix.copy(), poly_order=order)
        res.append(F * d)
    gc.close(res[1], res[0], what="FilterLO order %d" % order)
    # samples outside subscans, and (order > 0) flagged samples, are exactly zero
    inside = np.zeros(nt, dtype=bool)
    off = 0
    for L, S, ns, nb in zip(subs, tst, nsamples, nbolos):
        for b in range(nb):
            for ln, st in zip(L, S):
                inside[off + b * ns + st: off + b * ns + st + ln] = True
        off += ns * nb
    assert np.all(res[1][~inside] == 0.0)
    if order > 0:
        assert np.all(res[1][pix < 0] == 0.0)


def test_offset_filter_staged_equals_first_kernel(cm, variant):
    """poly_order = 0 through the shared-memory-staged kernel == cm2_filter_offset_apply (bit for bit
    is not promised: the partial sums are grouped differently)."""
    from cosmomap2_b200 import linearoperators as lo
    rng = np.random.default_rng(7)
    ns, nb = 50000, 4
    L, S = _subscan_table(rng, ns, 20)
    nt = ns * nb
    pix = rng.integers(0, 1000, size=nt).astype(np.int64)
    pix[rng.random(nt) < 0.05] = -1
    pix[S[3]:S[3] + L[3]] = -1
    d = rng.standard_normal(nt) + 5.0
    F = cm.FilterLO(nt, [L, S], ns, nb, pix)
    assert lo.FILTER_STAGED
    a = F * d
    lo.FILTER_STAGED = False
    try:
        b = F * d
    finally:
        lo.FILTER_STAGED = True
    gc.close(a, b, rtol=1e-13, what="staged offset filter vs first kernel")


def test_filter_unsorted_table_and_long_subscans(cm):
    """(1) a subscan table in reverse order takes the memset + sparse-write path; (2) subscans longer
    than the 24 000-sample shared-memory window re-read their tail from global memory."""
    import oracle
    rng = np.random.default_rng(11)
    ns, nb = 70000, 2
    L = np.array([30000, 26000, 5000], dtype=np.int64)
    S = np.array([100, 31000, 60000], dtype=np.int64)
    nt = ns * nb
    pix = rng.integers(0, 300, size=nt).astype(np.int64)
    pix[rng.random(nt) < 0.03] = -1
    pix[ns + S[1]: ns + S[1] + L[1]] = np.abs(pix[ns + S[1]: ns + S[1] + L[1]])      # one long subscan unflagged
    d = rng.standard_normal(nt) + np.linspace(-3, 3, nt) ** 2
    for order in (0, 2):
        ref = oracle.FilterLO(nt, [L, S], ns, nb, pix.copy(), poly_order=order) * d
        out = cm.FilterLO(nt, [L, S], ns, nb, pix.copy(), poly_order=order) * d
        gc.close(out, ref, what="long subscans, order %d" % order)
        Fr = cm.FilterLO(nt, [L[::-1].copy(), S[::-1].copy()], ns, nb, pix.copy(), poly_order=order)
        assert not Fr._sorted
        gc.close(Fr * d, ref, what="unsorted table, order %d" % order)


def test_legendre_filter_properties(cm, variant):
    """Size-independent properties at a size the oracle would not finish quickly: with flags the
    filter is an exact projector on each subscan (idempotent; polynomials of degree <= order are
    annihilated); linear."""
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(4000000, nside=256, ndet=16, nx=400, ny=200, samples_per_pixel=6.0, seed=2,
                               flag_turnarounds=True, with_data=False)
    rng = np.random.default_rng(3)
    pix = sc.pix.astype(np.int64)
    # at least one flag inside every subscan -> every subscan takes the QR (exact projector) branch
    for b in range(sc.ndet):
        pix[b * sc.ns + sc.sub_start + sc.sub_len // 2] = -1
    pix[rng.random(sc.nt) < 0.02] = -1
    order = 3
    F = cm.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix, poly_order=order)
    d1 = rng.standard_normal(sc.nt)
    d2 = rng.standard_normal(sc.nt)
    y1 = F * d1
    gc.close(F * y1, y1, rtol=1e-11, what="F F d = F d")
    gc.close(F * (2.0 * d1 - 0.5 * d2), 2.0 * y1 - 0.5 * (F * d2), rtol=1e-11, what="linearity")
    t = np.arange(sc.nt, dtype=np.float64) % sc.ns
    poly = 1.0 + 1e-3 * t - 2e-7 * t ** 2 + 1e-11 * t ** 3
    assert np.max(np.abs(F * poly)) <= 1e-8 * np.max(np.abs(poly))


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_fused_legendre_amatvec(cm, pol):
    """P^T F_K P as one kernel == the three-operator chain == the oracle's chain, and symmetric."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    from cosmomap2_b200 import synthetic
    sc = synthetic.raster_scan(300000, nside=64, ndet=6, nx=90, ny=50, samples_per_pixel=6.0, seed=5,
                               flag_turnarounds=True)
    rng = np.random.default_rng(9)
    res = {}
    for name, impl in (("oracle", oracle), ("gpu", cm)):
        pix = sc.pix.astype(np.int64)
        # flags only in the first half of every detector's timeline: the subscans of the second half
        # have no flagged sample and take the literal-sum branch, the others the QR branch
        flag = np.random.default_rng(1).random(sc.nt) < 0.03
        flag &= (np.arange(sc.nt) % sc.ns) < sc.ns // 2
        pix[flag] = -1
        # the leading 85 % of three subscans of detector 1 flagged: the unflagged samples cluster at the
        # end, the whole-subscan basis becomes ill-conditioned at the higher orders (local-basis path)
        for ks in (2, 3, 5):
            a0 = sc.ns + int(sc.sub_start[ks])
            pix[a0:a0 + int(0.85 * sc.sub_len[ks])] = -1
        pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        res[name] = dict(P=P, pix=pix, npix=npix)
    npix = res["gpu"]["npix"]
    assert npix == res["oracle"]["npix"]
    x = rng.standard_normal(pol * npix)
    z = rng.standard_normal(pol * npix)
    for order in (1, 2, 3, 4, 5):
        out = {}
        for name, impl in (("oracle", oracle), ("gpu", cm)):
            P, pix = res[name]["P"], res[name]["pix"]
            F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix, poly_order=order)
            A = P.T * F * P
            out[name] = A * x
            if name == "gpu":
                fused = A * x
                lo.fusion_enabled = False
                try:
                    chain = (P.T * F * P) * x
                finally:
                    lo.fusion_enabled = True
                gc.close(fused, chain, what="fused vs chain, order %d" % order)
                if order <= 4:
                    assert lo._FusedPolyFilterA.supported(P, F)
                # with flags in every subscan F is an orthogonal projector on the unflagged samples,
                # so A is symmetric; unflagged subscans use the reference's non-orthogonal sum, still
                # symmetric (sum_k b_k b_k^T)
                Az = A * z
                assert abs(np.dot(z, out[name]) - np.dot(x, Az)) <= 1e-10 * np.linalg.norm(Az) * np.linalg.norm(x)
        gc.close(out["gpu"], out["oracle"], what="P^T F_%d P x" % order)


def test_ground_filter_against_oracle(cm):
    import oracle
    rng = np.random.default_rng(13)
    nt = 500003
    az = (np.arange(nt) * 0.013) % 200.0                  # slow azimuth ramp: long runs per ground bin
    ground = np.floor(az).astype(np.int64)
    ground[ground == 17] = 18                             # a bin nobody hits (counts = 0)
    ground[rng.random(nt) < 0.04] = -1
    v = rng.standard_normal(nt) + 0.01 * ground
    Go = oracle.GroundFilterLO(ground.copy())
    Gg = cm.GroundFilterLO(ground.copy())
    assert Gg.nbins == Go.nbins and Gg.n == Go.n
    gc.close(Gg * v, Go * v, what="GroundFilterLO v")
    gc.close(Gg.Pg * v, Go.Pg * v, what="G (G^T G)^-1 G^T v")
    y = Gg * v
    gc.close(Gg * y, y, rtol=1e-12, what="idempotent")
    assert np.array_equal((Gg * v)[ground < 0], v[ground < 0])
    # random bins (no runs) and all-flagged input
    g2 = rng.integers(0, 1000, size=100000).astype(np.int64)
    v2 = rng.standard_normal(100000)
    gc.close(cm.GroundFilterLO(g2.copy()) * v2, oracle.GroundFilterLO(g2.copy()) * v2, what="random bins")


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_reorganize_map(cm, pol):
    import oracle
    import torch
    rng = np.random.default_rng(17)
    nside, npix = 64, 5000
    obspix = np.sort(rng.choice(12 * nside * nside, npix, replace=False))
    m = rng.standard_normal(pol * npix)
    ref = oracle.reorganize_map(m, obspix, npix, nside, pol)
    out = cm.reorganize_map(m, obspix, npix, nside, pol)
    assert len(out) == pol
    for a, b in zip(out, ref):
        assert np.array_equal(a, b)                       # a permutation: bit-exact
    dev = cm.reorganize_map(torch.from_numpy(m).cuda(), obspix, npix, nside, pol)
    assert all(t.is_cuda for t in dev) and np.array_equal(dev[0].cpu().numpy(), ref[0])
    with pytest.raises(IndexError):
        cm.reorganize_map(m, obspix + 12 * nside * nside, npix, nside, pol)


@pytest.mark.parametrize("name,pixscale,pol", [("testcase_block_diag_4.hdf5", 1, 1), ("testcase_block_diag_3.hdf5", 3, 3),
                                               ("testcase_block_diag_3.hdf5", 3, 2)])
def test_reference_hdf5_fixture_through_the_solve(cm, name, pixscale, pol):
    """The reference's own test-case files (h5py, system_setup(nt=100, npix=15, nb=2)) read by
    cosmomap2_b200.IOfiles and solved on the GPU and by the oracle: white and Toeplitz noise."""
    import os
    import scipy.sparse.linalg as spla
    import oracle
    det, pix_file, phi, weight = cm.read_from_hdf5(os.path.join(gc.GOLDEN, name))
    nb, nt = 2, len(det)
    res = {}
    for label, impl, solver in (("oracle", oracle, spla.cg), ("gpu", cm, cm.cg)):
        pix = (pix_file // pixscale).astype(np.int64)
        t = np.asarray(weight, dtype=np.float64).reshape(nb, -1)
        N = impl.BlockLO(nt // nb, t[:, 0].copy(), offdiag=False)
        pts = impl.ProcessTimeSamples(pix, 15, pol=pol, phi=phi, w=N.diag)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, nt, pix, pol=pol, angle_processed=pts)
        Mbd = impl.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
        b = P.T * (N * det)
        x, info = solver(P.T * N * P, b, M=Mbd, rtol=1e-12, maxiter=50)
        out = dict(npix=npix, pix=pix, b=b, x=x, info=info)
        if t.shape[1] > 1:                                 # the (nb, bandsize) noise values as Toeplitz bands
            band = t.copy()
            band[:, 0] += 1.0                              # diagonally dominant, as noise_val intends
            NT = impl.BlockLO(nt // nb, [r for r in band], offdiag=True)
            out["toep"] = P.T * (NT * (P * x))
        res[label] = out
    o, g = res["oracle"], res["gpu"]
    assert o["npix"] == g["npix"] and o["info"] == 0 and g["info"] == 0
    gc.exact(g["pix"], o["pix"], "relabelled pixels of the fixture")
    gc.close(g["b"], o["b"], what="b")
    gc.close(g["x"], o["x"], rtol=1e-10, what="x")
    if "toep" in o:
        gc.close(g["toep"], o["toep"], what="P^T N_toeplitz P x")


@pytest.mark.parametrize("pol,nt,npix,nb,ncv,tol", [(1, 500, 40, 1, 15, 1e-10), (3, 4000, 20, 1, 15, 1e-5),
                                                    (2, 1000, 20, 2, 50, 1e-4), (3, 60000, 400, 4, 40, 1e-8)])
def test_device_eigsh_equals_arpack(cm, pol, nt, npix, nb, ncv, tol):
    """cosmomap2_b200.eigsh (thick-restart Lanczos on device vectors) against SciPy's ARPACK on the
    oracle's operators, for the calls the reference's tests make (tests/test_2level_preconditioner.py:33,
    tests/test_coarse_operator.py:16, tests/test_deflation_operator.py:15; ncv > n is clamped like SciPy
    does) and a larger system: eigenvalues, the eigenspace, B-orthonormality, and the two-level
    identities M2 A z = z, R A z = 0 with the device-built Z."""
    import scipy.sparse.linalg as spla
    import oracle
    import reference_suite as rs
    no, Po, No, Mo, Bo, Ao, bo = rs._deflation_system(oracle, nt, npix, nb, pol, seed=70 + pol)
    ng, Pg, Ng, Mg, Bg, Ag, bg = rs._deflation_system(cm, nt, npix, nb, pol, seed=70 + pol)
    assert no == ng
    n = pol * ng
    w0, Z0 = spla.eigsh(Ao, M=Bo, Minv=Mo, k=5, v0=np.ones(n), which="SM", ncv=ncv, tol=tol)
    w1, Z1 = cm.eigsh(Ag, M=Bg, Minv=Mg, k=5, v0=np.ones(n), which="SM", ncv=ncv, tol=tol)
    order = np.argsort(w0)
    w0, Z0 = w0[order], Z0[:, order]
    assert np.allclose(w1, w0, rtol=max(100 * tol, 1e-9), atol=1e-12), (w0, w1)
    BZ1 = np.column_stack([Bo * Z1[:, i] for i in range(5)])
    assert np.abs(Z1.T.dot(BZ1) - np.eye(5)).max() < 1e-10             # Z^T B Z = I
    for i in range(5):                                                 # A z = lambda B z
        assert np.linalg.norm(Ao * Z1[:, i] - w1[i] * BZ1[:, i]) <= max(10 * tol, 1e-9) * np.linalg.norm(BZ1[:, i])
    # same eigenspace as ARPACK's (eigenvalues may be clustered: compare the projector, not the vectors)
    BZ0 = np.column_stack([Bo * Z0[:, i] for i in range(5)])
    assert np.linalg.norm(Z0 - Z1.dot(Z1.T.dot(BZ0))) <= max(1e4 * tol, 1e-8) * np.linalg.norm(Z0)
    # two-level preconditioner from the device-built Z (tests/test_2level_preconditioner.py:38-48)
    Az = np.column_stack([Ag * Z1[:, i] for i in range(5)])
    E = cm.CoarseLO(Z1, Az, 5)
    Zd = cm.DeflationLO(Z1)
    R = cm.lp.IdentityOperator(n) - Ag * Zd * E * Zd.T
    M2 = Mg * R + Zd * E * Zd.T
    for i in range(5):
        assert np.allclose(M2 * (Ag * Z1[:, i]), Z1[:, i], atol=1e-8)
        assert cm.norm2(R * (Ag * Z1[:, i])) <= 1e-9
    # other selections, without eigenvectors
    wl = cm.eigsh(Ag, M=Bg, Minv=Mg, k=3, which="LA", ncv=min(ncv, 20), tol=1e-10, v0=np.ones(n), return_eigenvectors=False)
    wl0 = spla.eigsh(Ao, M=Bo, Minv=Mo, k=3, which="LA", ncv=min(ncv, 20), tol=1e-10, v0=np.ones(n), return_eigenvectors=False)
    assert np.allclose(np.sort(wl), np.sort(wl0), rtol=1e-8)
    with pytest.raises(ValueError):
        cm.eigsh(Ag, M=Bg, k=3)


@pytest.mark.parametrize("nt", [1003, 1002, 1001, 1000, 7])
def test_subscan_filter_last_samples_of_the_tod(cm, variant, nt):
    """The last subscan ends at the very end of a TOD whose length is not a multiple of 4: the bulk
    copies stop at nt & ~3 and the remaining <= 3 samples are read directly; also subscans of 1-3
    samples and a single-subscan table."""
    import oracle
    rng = np.random.default_rng(nt)
    if nt > 100:
        L = np.array([3, 1, 200, 2, nt - 700], dtype=np.int64)
        S = np.array([0, 5, 10, 400, 700], dtype=np.int64)
    else:
        L, S = np.array([nt], dtype=np.int64), np.array([0], dtype=np.int64)
    pix = rng.integers(0, 9, size=nt).astype(np.int64)
    pix[rng.random(nt) < 0.1] = -1
    pix[-1] = 4
    d = rng.standard_normal(nt) + 2.0
    for order in (0, 1, 2):
        ref = oracle.FilterLO(nt, [L, S], nt, 1, pix.copy(), poly_order=order) * d
        out = cm.FilterLO(nt, [L, S], nt, 1, pix.copy(), poly_order=order) * d
        gc.close(out, ref, what="order %d, nt = %d" % (order, nt))


def test_ground_filter_edge_cases(cm):
    """All samples flagged (no bin at all): the filter is the identity; a single bin: the mean of the
    unflagged samples is removed from them."""
    rng = np.random.default_rng(2)
    v = rng.standard_normal(1001)
    g = np.full(1001, -1, dtype=np.int64)
    G = cm.GroundFilterLO(g.copy())
    assert G.nbins == 0 and np.array_equal(G * v, v)
    g[::3] = 0
    out = cm.GroundFilterLO(g.copy()) * v
    assert np.array_equal(out[g < 0], v[g < 0])
    assert np.allclose(out[g == 0], v[g == 0] - v[g == 0].mean(), rtol=0, atol=1e-14)
