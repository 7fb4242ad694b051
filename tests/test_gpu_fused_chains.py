"""GPU parity of the fused chains added after the first set: P^T T P for a short Toeplitz band in one
TOD pass (cm2_amatvec_toeplitz) and F P for the offset filter in one pass (cm2_pointing_filter_mu),
against the oracle and against the unfused chain of the same operators."""
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
    from cosmomap2_b200 import linearoperators as lo
    old = lo.FUSE_TOEPLITZ_A, lo.FUSE_FILTER_P
    lo.FUSE_TOEPLITZ_A = lo.FUSE_FILTER_P = True
    yield cosmomap2_b200
    lo.FUSE_TOEPLITZ_A, lo.FUSE_FILTER_P = old


def _raster(nt=300000, ndet=6, seed=3, **kw):
    from cosmomap2_b200 import synthetic
    return synthetic.raster_scan(nt, nside=64, ndet=ndet, nx=90, ny=50, samples_per_pixel=6.0, seed=seed, **kw)


def _bands(nblocks, nband, seed):
    rng = np.random.default_rng(seed)
    t = []
    for _ in range(nblocks):
        a = rng.standard_normal(nband) * 0.4 ** np.arange(nband)
        a[0] = 1.0 + rng.random()
        t.append(a)
    return t


def _setup(impl, sc, pol, t, blocksize=None):
    pix = sc.pix.astype(np.int64)
    N = impl.BlockLO(sc.ns if blocksize is None else blocksize, t, offdiag=True)
    pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi)
    npix = pts.get_new_pixel[0]
    P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
    return pix, N, P, npix


@pytest.mark.parametrize("pol", [1, 2, 3])
@pytest.mark.parametrize("nband", [1, 2, 3, 5, 8, 9])
def test_fused_toeplitz_amatvec(cm, pol, nband):
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(flag_turnarounds=True)
    sc.pix[np.random.default_rng(1).random(sc.nt) < 0.02] = -1
    t = _bands(sc.ndet, nband, seed=nband)
    _, No, Po, npix_o = _setup(oracle, sc, pol, t)
    _, Ng, Pg, npix = _setup(cm, sc, pol, t)
    assert npix == npix_o
    x = np.random.default_rng(2).standard_normal(pol * npix)
    ref = Po.T * (No * (Po * x))
    A = Pg.T * Ng * Pg
    y = A * x
    assert [type(f) for f in A.planned()] == [lo._FusedToeplitzA]
    gc.close(y, ref, what="fused P^T T P, pol=%d nband=%d" % (pol, nband))
    gc.close(y, Pg.T * (Ng * (Pg * x)), rtol=1e-12, what="fused == chain")
    # symmetric operator
    z = np.random.default_rng(3).standard_normal(pol * npix)
    assert abs(np.dot(z, y) - np.dot(x, A * z)) <= 1e-10 * abs(np.dot(z, y))


def test_fused_toeplitz_band_wider_than_the_fused_kernel_takes_the_chain(cm):
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(nt=60000, ndet=3)
    t = _bands(sc.ndet, 10, seed=5)
    _, No, Po, npix = _setup(oracle, sc, 3, t)
    _, Ng, Pg, _ = _setup(cm, sc, 3, t)
    x = np.random.default_rng(2).standard_normal(3 * npix)
    A = Pg.T * Ng * Pg
    gc.close(A * x, Po.T * (No * (Po * x)))
    assert not any(isinstance(f, lo._FusedToeplitzA) for f in A.planned())


@pytest.mark.parametrize("pol", [1, 3])
def test_fused_toeplitz_ragged_blocks_and_random_pointing(cm, pol):
    """Blocks of any size (shorter than the band, not multiples of the 8-sample chunk, boundaries inside
    a lane's chunk and at tile edges), TOD lengths around the tile sizes, run-free pointing."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    rng = np.random.default_rng(11)
    cases = [([1], 1), ([7], 3), ([9], 9), ([255], 2), ([257], 5), ([240, 240, 240], 3), ([224, 32, 1, 3, 500], 9),
             ([3, 2, 1, 1, 5], 4), ([1000, 999, 1001], 8), ([5000] * 4, 6), ([131, 4097, 77, 2048], 7)]
    for sizes, nband in cases:
        nt = sum(sizes)
        npix_full = 13
        pix = rng.integers(0, npix_full, nt)
        if nt > 300:
            pix[::7] = -1
            pix = np.where(rng.random(nt) < 0.5, np.sort(pix), pix)      # runs and run-free stretches
        phi = rng.random(nt) * np.pi
        t = _bands(len(sizes), nband, seed=nt)
        x = None
        res = []
        for impl in (oracle, cm):
            p = pix.astype(np.int64).copy()
            pts = impl.ProcessTimeSamples(p, npix_full, pol=pol, phi=phi)
            npix = pts.get_new_pixel[0]
            if npix == 0:
                res.append(None)
                continue
            P = impl.SparseLO(npix, nt, p, pol=pol, angle_processed=pts)
            N = impl.BlockLO(sizes, t, offdiag=True)
            if x is None:
                x = rng.standard_normal(pol * npix)
            if impl is oracle:
                res.append(P.T * (N * (P * x)))
            else:
                A = P.T * N * P
                res.append(A * x)
                assert [type(f) for f in A.planned()] == [lo._FusedToeplitzA]
        if res[0] is not None:
            gc.close(res[1], res[0], what="sizes=%s nband=%d" % (sizes, nband))


def test_fused_toeplitz_error_paths(cm):
    import torch
    from cosmomap2_b200 import _cabi, _device as dv
    pix = torch.zeros(8, dtype=torch.int32, device="cuda")
    band = torch.ones(12, dtype=torch.float64, device="cuda")
    x = torch.ones(1, dtype=torch.float64, device="cuda")
    y = torch.ones(1, dtype=torch.float64, device="cuda")
    with pytest.raises(_cabi.Cm2Error):                 # band wider than the kernel supports
        dv.call("cm2_amatvec_toeplitz", dv.ptr(pix), None, None, 8, 1, dv.ptr(band), 12, 1, 8, None, dv.ptr(x), dv.ptr(y), 1, None)
    with pytest.raises(_cabi.Cm2Error):                 # no band
        dv.call("cm2_amatvec_toeplitz", dv.ptr(pix), None, None, 8, 1, None, 1, 1, 8, None, dv.ptr(x), dv.ptr(y), 1, None)
    # empty TOD: y = 0
    dv.call("cm2_amatvec_toeplitz", dv.ptr(pix), None, None, 0, 1, dv.ptr(band), 3, 1, 8, None, dv.ptr(x), dv.ptr(y), 1, None)
    assert float(y.item()) == 0.0


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_fused_filter_pointing(cm, pol):
    """F P in one pass == oracle F (P x), flagged samples inside subscans (they get -mean), a fully
    flagged subscan (stays 0) and the gaps (0) included; and the chain P.T*F*N*F*P built on it."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(nt=300000, ndet=6, seed=8, flag_turnarounds=True)
    rng = np.random.default_rng(3)
    sc.pix[rng.random(sc.nt) < 0.02] = -1
    k = 5                                               # one subscan of detector 2 fully flagged
    a = 2 * sc.ns + int(sc.sub_start[k])
    sc.pix[a:a + int(sc.sub_len[k])] = -1
    for nband in (3, 60, 700):                          # fused-band range / FFT range (one / several windows apart)
        t = _bands(sc.ndet, nband, seed=nband)
        out = {}
        for name, impl in (("oracle", oracle), ("gpu", cm)):
            pix, N, P, npix = _setup(impl, sc, pol, t)
            F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix)
            x = np.random.default_rng(4).standard_normal(pol * npix)
            if impl is oracle:
                d = F * (P * x)
                out[name] = (d, P.T * (F * (N * d)))
            else:
                FP = F * P
                d = FP * x
                assert [type(f) for f in FP.planned()] == [lo._FusedFilterP]
                assert FP.planned()[0]._runs
                A = P.T * F * N * F * P
                y = A * x
                assert isinstance(A.planned()[-1], lo._FusedFilterP)
                out[name] = (d, y)
        gc.close(out["gpu"][0], out["oracle"][0], what="F P x")
        gc.close(out["gpu"][1], out["oracle"][1], what="P^T F N F P x, nband=%d" % nband)


def test_fused_filter_pointing_falls_back_without_runs(cm):
    """Random pointing has no runs: the run table is refused and F P runs as two operators."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    rng = np.random.default_rng(5)
    nt, npix_full = 40000, 50
    pix0 = rng.integers(0, npix_full, nt)
    sub_len = np.full(20, 1500)
    sub_start = np.arange(20) * 2000 + 100
    out = []
    for impl in (oracle, cm):
        pix = pix0.astype(np.int64).copy()
        pts = impl.ProcessTimeSamples(pix, npix_full, pol=1)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, nt, pix, pol=1, angle_processed=pts)
        F = impl.FilterLO(nt, [sub_len, sub_start], nt, 1, pix)
        x = np.random.default_rng(6).standard_normal(npix)
        if impl is oracle:
            out.append(F * (P * x))
        else:
            FP = F * P
            out.append(FP * x)
            assert isinstance(FP.planned()[0], lo._FusedFilterP) and FP.planned()[0]._runs is False
    gc.close(out[1], out[0])


# ---- the run-table Legendre A-matvec (promoted in round 2 after its first B200 run: order 1 0.61 vs 0.79 ms) ----
@pytest.mark.parametrize("pol", [1, 3])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_poly_run_table_path(cm, pol, order):
    """P^T F_K P through the Legendre run table (one TOD pass) == the per-subscan kernel == oracle, with
    scattered flags, a subscan that keeps only its first tenth (ill-conditioned at every order: stays with the
    per-subscan kernel) and a subscan with fewer unflagged samples than the order (skipped by the reference)."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(nt=300000, ndet=6, seed=8, flag_turnarounds=True)
    rng = np.random.default_rng(3)
    sc.pix[rng.random(sc.nt) < 0.02] = -1
    a = 1 * sc.ns + int(sc.sub_start[3])
    sc.pix[a + int(sc.sub_len[3]) // 10:a + int(sc.sub_len[3])] = -1
    a = 4 * sc.ns + int(sc.sub_start[7])
    keep = sc.pix[a]
    sc.pix[a:a + int(sc.sub_len[7])] = -1
    sc.pix[a] = keep
    res = {}
    for name, impl, table in (("oracle", oracle, None), ("table", cm, True), ("subscan", cm, False)):
        pix = sc.pix.astype(np.int64)
        pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix, poly_order=order)
        x = np.random.default_rng(4).standard_normal(pol * npix)
        if table is None:
            res[name] = P.T * (F * (P * x))
            continue
        old = lo.FILTER_POLY_RUN_TABLE
        lo.FILTER_POLY_RUN_TABLE = table
        try:
            A = P.T * F * P
            res[name] = A * x
            fused = [f for f in A.planned() if isinstance(f, lo._FusedPolyFilterA)][0]
            assert bool(fused._runs) == table
            if table:
                assert fused._runs["nhard"] >= 1
        finally:
            lo.FILTER_POLY_RUN_TABLE = old
    gc.close(res["subscan"], res["oracle"], what="per-subscan kernel")
    gc.close(res["table"], res["oracle"], what="run-table path")


@pytest.mark.parametrize("pol", [1, 3])
@pytest.mark.parametrize("kind", ["white", "offset", "legendre"])
def test_stream_interleaved_tile_order_equals_time_order(cm, pol, kind):
    """The single-pass A-matvecs walk the TOD detector-interleaved when the map is larger than L2
    (``nstreams`` in include/cosmomap2_b200.h): the same sum in another order of the atomic adds.  Forced here
    on a small scan -- detector timelines that are not a multiple of the 256-sample tile, a last stream that is
    short -- against the time-ordered kernel and the oracle."""
    import oracle
    from cosmomap2_b200 import linearoperators as lo
    sc = _raster(nt=7 * 43211, ndet=7, seed=5, flag_turnarounds=True)
    sc.pix[np.random.default_rng(3).random(sc.nt) < 0.01] = -1
    res = {}
    for label, impl in (("oracle", oracle), ("time", cm), ("streams", cm)):
        pix = sc.pix.astype(np.int64)
        pts = impl.ProcessTimeSamples(pix, sc.npix_full, pol=pol, phi=sc.phi)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, sc.nt, pix, pol=pol, angle_processed=pts)
        if kind == "white":
            A = P.T * impl.BlockLO(sc.ns, sc.weights) * P
        else:
            F = impl.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, pix, poly_order=0 if kind == "offset" else 2)
            A = P.T * F * P
        x = np.random.default_rng(4).standard_normal(pol * npix)
        old = lo.TOD_INTERLEAVE_MIN_MAP_BYTES
        lo.TOD_INTERLEAVE_MIN_MAP_BYTES = 0 if label == "streams" else 1e30
        try:
            res[label] = A * x
            if impl is cm:
                fused = [f for f in A.planned() if isinstance(f, (lo._FusedWhiteA, lo._FusedFilterA, lo._FusedPolyFilterA))]
                assert len(fused) == 1
                assert lo._tod_streams(P, sc.ns) == (sc.ndet if label == "streams" else 1)
        finally:
            lo.TOD_INTERLEAVE_MIN_MAP_BYTES = old
    gc.close(res["streams"], res["time"], rtol=1e-13, what="interleaved vs time order")
    gc.close(res["streams"], res["oracle"], rtol=1e-10, what="interleaved vs oracle")


@pytest.mark.parametrize("pol", [1, 2, 3])
@pytest.mark.parametrize("pattern", ["spp3", "spp1", "random"])
def test_white_amatvec_scatter_modes(cm, pol, pattern):
    """The fused white A-matvec chooses its scatter from the pointing (linearoperators._FusedWhiteA): staged
    through shared memory for short runs, a pixel-sorted copy of the pointing for run-free pointing (the
    reference tests' random pointing, utilities/utilities_functions.py:111-122).  Every mode against the oracle
    and against the register path, with flagged samples and unequal block weights."""
    import oracle
    from cosmomap2_b200 import synthetic, linearoperators as lo
    rng = np.random.default_rng(11)
    if pattern == "random":
        ndet, ns, nside = 5, 30011, 16
        nt = ndet * ns
        pix0 = rng.integers(0, 12 * nside * nside, nt).astype(np.int32)
        phi = rng.uniform(0, np.pi, nt)
        weights = rng.uniform(0.5, 1.5, ndet)
        npix_full = 12 * nside * nside
    else:
        sc = synthetic.raster_scan(6 * 40003, nside=64, ndet=6, nx=90, ny=50,
                                   samples_per_pixel=4.0 if pattern == "spp3" else 1.0, seed=7, flag_turnarounds=True)
        pix0, phi, weights, ns, nt, npix_full = sc.pix.copy(), sc.phi, sc.weights, sc.ns, sc.nt, sc.npix_full
    pix0[rng.random(nt) < 0.02] = -1
    res, modes = {}, {}
    for label, impl in (("oracle", oracle), ("auto", cm), ("registers", cm)):
        pix = pix0.astype(np.int64)
        N = impl.BlockLO(ns, weights)
        pts = impl.ProcessTimeSamples(pix, npix_full, pol=pol, phi=phi, w=N.diag)
        npix = pts.get_new_pixel[0]
        P = impl.SparseLO(npix, nt, pix, pol=pol, angle_processed=pts)
        A = P.T * N * P
        x = np.random.default_rng(4).standard_normal(pol * npix)
        old = lo.WHITE_SORT_BELOW_RUN, lo.WHITE_STAGE_RUN_RANGE, lo.WHITE_STAGE_MIN_CONTIGUITY
        lo.WHITE_STAGE_MIN_CONTIGUITY = 0.5              # the 2 % random flags break some of the +-1 steps
        if label == "registers":
            lo.WHITE_SORT_BELOW_RUN, lo.WHITE_STAGE_RUN_RANGE = 0.0, (0.0, 0.0)
        try:
            res[label] = A * x
            if impl is cm:
                fused = [f for f in A.planned() if isinstance(f, lo._FusedWhiteA)]
                assert len(fused) == 1
                modes[label] = fused[0]._mode
        finally:
            lo.WHITE_SORT_BELOW_RUN, lo.WHITE_STAGE_RUN_RANGE, lo.WHITE_STAGE_MIN_CONTIGUITY = old
    assert modes["registers"] == "registers"
    assert modes["auto"] == {"spp3": "staged", "spp1": "sorted", "random": "sorted"}[pattern]
    gc.close(res["auto"], res["registers"], rtol=1e-13, what="%s vs register path" % modes["auto"])
    gc.close(res["auto"], res["oracle"], rtol=1e-10, what="%s vs oracle" % modes["auto"])
