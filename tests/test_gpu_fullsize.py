"""GPU parity at BASELINE.json's full single-GPU size (configs[1]: 1e8 samples, IQU nside 512) through
size-independent properties -- the oracle cannot run this size in seconds, so the checks are the
identities the reference's own tests use (tests/test_matrix_vector_product.py:9-23, 65-94;
tests/test_block_diagonal_operator.py:39-64) plus adjointness, fused == unfused and idempotence of the
time-domain filters."""
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


@pytest.fixture(scope="module")
def c2(cm):
    import torch
    from cosmomap2_b200 import synthetic
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs 40 GB of device memory")
    sc = synthetic.config_c2(with_data=False)
    pol = 3
    nflag_in = int(np.count_nonzero(sc.pix < 0))
    N = cm.BlockLO(sc.ns, sc.weights)
    pts = cm.ProcessTimeSamples(sc.pix, sc.npix_full, pol=pol, phi=sc.phi, w=N.diag)
    npix = pts.get_new_pixel[0]
    P = cm.SparseLO(npix, sc.nt, sc.pix, pol=pol, angle_processed=pts)
    Mbd = cm.BlockDiagonalPreconditionerLO(pts, npix, pol=pol)
    return dict(sc=sc, N=N, pts=pts, npix=npix, P=P, Mbd=Mbd, pol=pol, nflag_in=nflag_in)


def _dev(a):
    from cosmomap2_b200 import _device as dv
    return dv.to_dev_f64(a)


def test_fullsize_hit_counts_and_relabelling(cm, c2):
    sc, P, npix = c2["sc"], c2["P"], c2["npix"]
    assert sc.nt == 100000000
    hits = P.hits()
    nflag = int(np.count_nonzero(sc.pix < 0))            # pixs was relabelled in place
    assert int(hits.sum()) == sc.nt - nflag               # bit-exact integer bookkeeping
    assert sc.pix.max() == npix - 1 and sc.pix.min() >= -1
    assert nflag >= c2["nflag_in"]
    # P^T P 1 == counts, exactly (integers in fp64)
    ones = np.ones(sc.nt)
    P1 = cm.SparseLO(npix, sc.nt, sc.pix)
    assert np.array_equal(P1.T * ones, hits.astype(np.float64))


def test_fullsize_mbd_a_is_identity_and_a_symmetric(cm, c2):
    import torch
    P, N, Mbd, npix, pol = c2["P"], c2["N"], c2["Mbd"], c2["npix"], c2["pol"]
    rng = np.random.default_rng(0)
    x = _dev(rng.standard_normal(pol * npix))
    y = _dev(rng.standard_normal(pol * npix))
    A = P.T * N * P
    Ax = A * x
    err = (Mbd * Ax - x).abs().max().item() / x.abs().max().item()
    assert err < 1e-9, err                                 # white noise, w = N.diag: M_BD A = I
    Ay = A * y
    sym = abs(torch.dot(y, Ax).item() - torch.dot(x, Ay).item()) / (Ax.norm().item() * y.norm().item())
    assert sym < 1e-12, sym
    assert torch.dot(x, Ax).item() > 0.0
    # fused kernel == the three-operator chain with its TOD temporaries
    from cosmomap2_b200 import linearoperators as lo
    lo.fusion_enabled = False
    try:
        chain = (P.T * N * P) * x
    finally:
        lo.fusion_enabled = True
    assert (chain - Ax).abs().max().item() <= 1e-11 * Ax.abs().max().item()
    # deterministic (pixel-sorted) transpose == atomic transpose
    d = N * (P * x)
    assert (_dev(P.rmult_sorted(d)) - P.T * d).abs().max().item() <= 1e-11 * Ax.abs().max().item()


def test_fullsize_adjointness_and_one_iteration_pcg(cm, c2):
    import torch
    sc, P, N, Mbd, npix, pol = c2["sc"], c2["P"], c2["N"], c2["Mbd"], c2["npix"], c2["pol"]
    rng = np.random.default_rng(1)
    x = _dev(rng.standard_normal(pol * npix))
    d = _dev(rng.standard_normal(sc.nt))
    lhs = torch.dot(P * x, d).item()                      # <P x, d> == <x, P^T d>
    rhs = torch.dot(x, P.T * d).item()
    assert abs(lhs - rhs) <= 1e-11 * (P * x).norm().item() * d.norm().item()
    b = P.T * (N * d)
    sol, info = cm.cg(P.T * N * P, b, M=Mbd, rtol=1e-10, maxiter=5)
    assert info == 0
    res = ((P.T * N * P) * sol - b).norm().item() / b.norm().item()
    assert res < 1e-10, res                               # converged after the first iteration (src/test_BD...:52)


@pytest.mark.parametrize("order", [0, 1, 3])
def test_fullsize_subscan_filters(cm, c2, order):
    """Offset and Legendre filters over 1e8 samples: linear, idempotent where the reference's filter is a
    projector (order 0 on unflagged data; Legendre with a flag in every subscan), zero in the gaps,
    and P^T F P (fused) == the chain."""
    import torch
    sc, P, npix, pol = c2["sc"], c2["P"], c2["npix"], c2["pol"]
    rng = np.random.default_rng(2)
    F = cm.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix, poly_order=order)
    d = _dev(rng.standard_normal(sc.nt))
    y = F * d
    if order == 0:
        assert (F * y - y).abs().max().item() <= 1e-12 * y.abs().max().item()
    x = _dev(rng.standard_normal(pol * npix))
    A = P.T * F * P
    Ax = A * x
    from cosmomap2_b200 import linearoperators as lo
    lo.fusion_enabled = False
    try:
        chain = (P.T * F * P) * x
    finally:
        lo.fusion_enabled = True
    assert (chain - Ax).abs().max().item() <= 1e-10 * Ax.abs().max().item()
    z = _dev(rng.standard_normal(pol * npix))
    sym = abs(torch.dot(z, Ax).item() - torch.dot(x, A * z).item()) / (Ax.norm().item() * z.norm().item())
    assert sym < 1e-11, sym
    # a map that is constant in I and zero in Q, U is filtered out entirely: P x is constant on every subscan
    mono = torch.zeros(pol * npix, dtype=torch.float64, device=x.device)
    mono[0::pol] = 3.0
    if order <= 1:      # from order 2 on the unflagged branch is the reference's non-orthogonal sum, not a projector
        assert (A * mono).abs().max().item() <= 1e-9 * 3.0 * (sc.nt / npix)


@pytest.mark.parametrize("nband", [1, 3, 9])
def test_fullsize_fused_toeplitz_amatvec(cm, c2, nband):
    """P^T T P in one TOD pass over 1e8 samples (64 noise blocks): equal to the chain P, T, P^T with its
    two TOD temporaries, symmetric, and -- with a one-coefficient band -- equal to the white-noise kernel."""
    import torch
    from cosmomap2_b200 import linearoperators as lo, synthetic
    sc, P, N, npix, pol = c2["sc"], c2["P"], c2["N"], c2["npix"], c2["pol"]
    rng = np.random.default_rng(3)
    x = _dev(rng.standard_normal(pol * npix))
    if nband == 1:
        bands = [[w] for w in sc.weights]
    else:
        bands = synthetic.toeplitz_bands(sc.ndet, nband, seed=nband)
    Nt = cm.BlockLO(sc.ns, bands, offdiag=True)
    A = P.T * Nt * P
    Ax = A * x
    assert [type(f) for f in A.planned()] == [lo._FusedToeplitzA]
    lo.fusion_enabled = False
    try:
        chain = (P.T * Nt * P) * x
    finally:
        lo.fusion_enabled = True
    assert (chain - Ax).abs().max().item() <= 1e-11 * Ax.abs().max().item()
    z = _dev(rng.standard_normal(pol * npix))
    sym = abs(torch.dot(z, Ax).item() - torch.dot(x, A * z).item()) / (Ax.norm().item() * z.norm().item())
    assert sym < 1e-12, sym
    if nband == 1:
        white = (P.T * N * P) * x
        assert (white - Ax).abs().max().item() <= 1e-12 * Ax.abs().max().item()


def test_fullsize_fused_filter_pointing(cm, c2):
    """F P in one pass over 1e8 samples == the chain; a map constant in I is removed inside every subscan."""
    import torch
    from cosmomap2_b200 import linearoperators as lo
    sc, P, npix, pol = c2["sc"], c2["P"], c2["npix"], c2["pol"]
    F = cm.FilterLO(sc.nt, [sc.sub_len, sc.sub_start], sc.ns, sc.ndet, sc.pix)
    x = _dev(np.random.default_rng(4).standard_normal(pol * npix))
    FP = F * P
    d = FP * x
    assert [type(f) for f in FP.planned()] == [lo._FusedFilterP] and FP.planned()[0]._runs
    chain = F * (P * x)
    assert (chain - d).abs().max().item() <= 1e-11 * chain.abs().max().item()
    mono = torch.zeros(pol * npix, dtype=torch.float64, device=x.device)
    mono[0::pol] = 3.0
    dm = FP * mono
    # unflagged samples: exactly filtered; flagged samples inside a subscan carry -mean = -3 (FilterLO.mult :165)
    pix = torch.as_tensor(sc.pix, device=x.device)
    assert dm[pix >= 0].abs().max().item() <= 1e-9
    assert dm.abs().max().item() <= 3.0 + 1e-9
