"""CPU tests of the product's host-side logic (no GPU, no kernels): subscan flattening, noise-block
partition, lazy block weights, the host-built FFT transfer-function tables (checked by emulating the
kernel's in-place bit-reversed algorithm in NumPy), the synthetic generators and the operator
algebra's planning of fused chains."""
import numpy as np
import pytest

import oracle
from cosmomap2_b200 import linearoperators as lo
from cosmomap2_b200 import synthetic
from cosmomap2_b200.process_ces import BlockWeights


def test_flatten_subscans_matches_reference_triple_loop():
    rng = np.random.default_rng(0)
    nsamples, nbolos = [400, 300, 250], [3, 2, 4]
    subs = [np.array([50, 60, 70]), np.array([40, 100]), np.array([30, 30, 30, 30])]
    tst = [np.array([5, 100, 250]), np.array([10, 120]), np.array([0, 50, 110, 200])]
    s, e = lo.flatten_subscans(subs, tst, nsamples, nbolos)
    ref_s, ref_e = [], []
    offset = 0
    for subsc, ts, ns, nb in zip(subs, tst, nsamples, nbolos):       # linearoperators.py:134-167
        for bolo in range(nb):
            for i, j in zip(subsc, ts):
                start = j + ns * bolo + offset
                ref_s.append(start)
                ref_e.append(start + i)
        offset += nb * ns
    assert np.array_equal(s, ref_s) and np.array_equal(e, ref_e)
    assert np.all(s[1:] >= e[:-1])          # sorted, non-overlapping: the run-table path applies
    # and the flattened filter equals the oracle's FilterLO
    nt = offset
    pix = rng.integers(0, 9, nt)
    pix[rng.random(nt) < 0.1] = -1
    d = rng.standard_normal(nt)
    F = oracle.FilterLO(nt, [subs, tst], nsamples, nbolos, pix)
    from oracle import cloops
    assert np.allclose(cloops.filter_offset(pix, d, s, e), F * d, rtol=0, atol=1e-13)


def test_block_starts_and_lazy_weights():
    st = lo._block_starts(100, 4)
    assert np.array_equal(st, [0, 100, 200, 300, 400])
    st = lo._block_starts([5, 7, 3], 3)
    assert np.array_equal(st, [0, 5, 12, 15])
    with pytest.raises(ValueError):
        lo._block_starts([5, 7], 3)
    w = BlockWeights([2.0, 3.0, 4.0], st)
    assert len(w) == 15 and w.shape == (15,) and w.equal_blocksize() == 0
    assert np.array_equal(np.asarray(w), np.repeat([2.0, 3.0, 4.0], [5, 7, 3]))
    assert np.array_equal(np.asarray(w), oracle.BlockLO([5, 7, 3], [2.0, 3.0, 4.0]).diag)
    assert BlockWeights([1.0, 2.0], [0, 8, 16]).equal_blocksize() == 8


def _emulate_fft_kernel(coef_block, a, v, log2m):
    """NumPy twin of k_toeplitz_fft for one noise block (in-place DIF -> tables by physical
    position -> in-place DIT), used to pin the host-built tables without a GPU."""
    M = 1 << log2m
    NF = 2 * M
    L = len(a)
    n = len(v)
    tw = np.exp(-2j * np.pi * np.arange(M // 2) / M)
    brev = np.array([int(format(k, "0%db" % log2m)[::-1], 2) for k in range(M)])
    S = NF - 2 * (L - 1)
    out = np.zeros(n)
    for win in range((n + S - 1) // S):
        j0 = win * S
        x = np.zeros(NF)
        t = j0 - (L - 1) + np.arange(NF)
        ok = (t >= 0) & (t < n)
        x[ok] = v[t[ok]]
        z = x[0::2] + 1j * x[1::2]
        for lm in range(log2m - 1, -1, -1):           # DIF, natural -> bit-reversed
            m = 1 << lm
            j = np.arange(M // 2)
            pos = j & (m - 1)
            i = ((j >> lm) << (lm + 1)) + pos
            A, B = z[i].copy(), z[i + m].copy()
            z[i] = A + B
            z[i + m] = (A - B) * tw[pos << (log2m - 1 - lm)]
        pk = np.arange(M)
        k = brev[pk]
        km = (M - k) & (M - 1)
        pm = brev[km]
        zk, zm = z[pk].copy(), z[pm].copy()
        E = 0.5 * (zk + np.conj(zm))
        O = (zk - np.conj(zm)) / 2j
        z = coef_block[0][pk] * E + coef_block[1][pk] * O      # W at physical position pk
        for lm in range(log2m):                        # DIT, bit-reversed -> natural
            m = 1 << lm
            j = np.arange(M // 2)
            pos = j & (m - 1)
            i = ((j >> lm) << (lm + 1)) + pos
            A, B = z[i].copy(), z[i + m] * np.conj(tw[pos << (log2m - 1 - lm)])
            z[i] = A + B
            z[i + m] = A - B
        zr = np.empty(NF)
        zr[0::2], zr[1::2] = z.real, z.imag
        idx = j0 + np.arange(S)
        keep = idx < n
        out[idx[keep]] = zr[L - 1 + np.arange(S)][keep]
    return out


@pytest.mark.parametrize("L", [1, 2, 9, 40])
def test_fft_transfer_tables_reproduce_the_toeplitz_product(L):
    rng = np.random.default_rng(L)
    log2m = 7
    a = rng.random(L)
    a[0] += 2.0
    v = rng.standard_normal(700)
    coef = lo.toeplitz_fft_tables(a[None, :], L, 1 << log2m)
    y = _emulate_fft_kernel(coef[0], a, v, log2m)
    ref = oracle.ToeplitzLO(a, len(v)) * v
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))


def _emulate_fft_pair_kernel(coef, a, v, log2m):
    """NumPy emulation of the 2-CTA (pair) mode of csrc/toeplitz_fft.cu on one block: windows of 4M samples, one
    2M-point packed transform split by a radix-2 stage -- CTA 0: u = z[:M] + z[M:] (even frequencies), CTA 1:
    v = (z[:M] - z[M:]) W^n (odd); transfer step inside each CTA (partner M - k' / M - 1 - k'); the halves of the
    window are U +- conj(W^n) V.  ``coef``: [2 (CTA)][2][M] as built by toeplitz_fft_tables(pair=True), entries at
    the bit-reversed position of k'."""
    M = 1 << log2m
    NF2, L, n = 4 * M, len(a), len(v)
    S = NF2 - 2 * (L - 1)
    brev = lo._bit_reverse(M)
    out = np.zeros(n)
    nn = np.arange(M)
    Wn = np.exp(-2j * np.pi * nn / (2 * M))
    kp = np.arange(M)
    for j0 in range(0, n, S):
        w0 = j0 - (L - 1)
        t = w0 + np.arange(NF2)
        x = np.where((t >= 0) & (t < n), v[np.clip(t, 0, n - 1)], 0.0)
        z = x[0::2] + 1j * x[1::2]
        halves = (z[:M] + z[M:], (z[:M] - z[M:]) * Wn)
        res = []
        for c in range(2):
            Z = np.fft.fft(halves[c])                              # natural order; the kernel keeps Z[k'] at brev(k')
            partner = np.conj(Z[(M - kp) % M]) if c == 0 else np.conj(Z[M - 1 - kp])
            E, O = 0.5 * (Z + partner), (Z - partner) / 2j
            C1, C2 = np.empty(M, complex), np.empty(M, complex)
            C1[kp] = coef[c, 0][brev[kp]]                           # table[brev(k')] holds frequency 2k' + c
            C2[kp] = coef[c, 1][brev[kp]]
            res.append(np.fft.ifft(C1 * E + C2 * O) * M)
        U, V = res
        w = np.concatenate([U + np.conj(Wn) * V, U - np.conj(Wn) * V])
        zr = np.empty(NF2)
        zr[0::2], zr[1::2] = w.real, w.imag
        idx = j0 + np.arange(S)
        keep = idx < n
        out[idx[keep]] = zr[L - 1 + np.arange(S)][keep]
    return out


@pytest.mark.parametrize("L", [1, 2, 9, 40, 100])
def test_fft_pair_tables_reproduce_the_toeplitz_product(L):
    rng = np.random.default_rng(L)
    log2m = 7
    a = rng.random(L)
    a[0] += 2.0
    v = rng.standard_normal(1900)
    coef = lo.toeplitz_fft_tables(a[None, :], L, 1 << log2m, pair=True)
    assert coef.shape == (1, 2, 2, 1 << log2m)
    y = _emulate_fft_pair_kernel(coef[0], a, v, log2m)
    ref = oracle.ToeplitzLO(a, len(v)) * v
    assert np.max(np.abs(y - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_synthetic_scan_properties():
    sc = synthetic.raster_scan(120000, nside=64, ndet=6, nx=50, ny=30, samples_per_pixel=7.0, seed=3,
                               flag_turnarounds=True)
    assert sc.nt == 120000 and len(sc.pix) == sc.nt and sc.pix.dtype == np.int32
    good = sc.pix >= 0
    assert 0.9 < good.mean() < 0.97                          # 5 % turnarounds flagged
    assert sc.pix[good].max() < sc.npix_full
    runs = np.count_nonzero(np.diff(sc.pix[good]) != 0)
    assert 5.0 < good.sum() / runs < 8.0                      # ~samples_per_pixel samples per pixel crossing
    # every unflagged sample lies inside a subscan of its detector, every flagged one outside
    inside = np.zeros(sc.nt, dtype=bool)
    s, e = lo.flatten_subscans([sc.sub_len], [sc.sub_start], [sc.ns], [sc.ndet])
    for a, b in zip(s, e):
        inside[a:b] = True
    assert np.all(inside[good])
    sc2 = synthetic.raster_scan(120000, nside=64, ndet=6, nx=50, ny=30, samples_per_pixel=7.0, seed=3,
                                flag_turnarounds=True)
    assert np.array_equal(sc.pix, sc2.pix) and np.array_equal(sc.phi, sc2.phi) and np.array_equal(sc.d, sc2.d)
    bands = synthetic.toeplitz_bands(3, 16)
    for a in bands:                                            # diagonally dominant -> SPD Toeplitz
        assert a[0] > 2 * np.abs(a[1:]).sum()


def test_small_utilities_of_the_reference_surface():
    """profile_run / output_profile (utilities_functions.py:65-89), rescalepixels (:91-96),
    subtract_offset (healpy_functions.py:146-158)."""
    import cosmomap2_b200 as cm
    pr = cm.profile_run()
    pr.enable()
    sum(range(100))
    pr.disable()
    cm.output_profile(pr)
    lo, shifted, hi = cm.rescalepixels(np.array([7, 9, 12]))
    assert (lo, hi) == (7, 12) and shifted.tolist() == [0, 2, 5]
    obs = np.array([1, 2, 3])
    m = [np.arange(10.0), np.ones(10)]
    cm.subtract_offset(m, obs, 3)
    assert abs(m[0][obs].mean()) < 1e-15 and m[0][0] == 0.0 and np.all(m[1][obs] == 0.0) and m[1][0] == 1.0
    one = np.arange(10.0)
    cm.subtract_offset(one, np.array([0, 9]), 1)
    assert one[0] == -4.5 and one[9] == 4.5 and one[5] == 5.0


def test_tile_tables_against_brute_force():
    """The per-tile subscan lookup the single-pass filter kernels use (linearoperators._tile_tables)."""
    import torch
    from cosmomap2_b200 import linearoperators as lo
    rng = np.random.default_rng(0)
    for trial in range(30):
        nseg = int(rng.integers(1, 12))
        lens = rng.integers(1, 900, nseg)
        gaps = rng.integers(0, 400, nseg + 1)
        start = np.cumsum(gaps[:-1]) + np.concatenate([[0], np.cumsum(lens[:-1])])
        end = start + lens
        nt = int(end[-1] + gaps[-1])
        tile_seg, tile_flag = lo._tile_tables(torch.as_tensor(start), torch.as_tensor(end), nseg, nt)
        ntiles = (nt + 255) // 256
        assert tile_seg.dtype == torch.int32 and tile_seg.numel() == ntiles and tile_flag.numel() == ntiles
        inside = np.zeros(nt, dtype=bool)
        for a, b in zip(start, end):
            inside[a:b] = True
        for i in range(ntiles):
            t0, t1 = 256 * i, min(256 * i + 256, nt)
            k = int(np.searchsorted(end, t0, side="right"))        # first segment ending beyond t0
            assert int(tile_seg[i]) == k
            if not inside[t0:t1].any():
                want = 0
            elif k < nseg and start[k] <= t0 and t1 <= end[k]:
                want = 1
            else:
                want = 2
            assert int(tile_flag[i]) == want, (trial, i)


def test_run_table_legendre_scheme_equals_polyfilter():
    """The arithmetic behind the single-TOD-pass Legendre A-matvec (csrc/filter_runs.cu: c = W S with S from
    a run table of Legendre-weighted sums, W = diag(1/||L_k||^2) without flags and the inverse Gram matrix
    otherwise, ill-conditioned subscans left to the per-subscan kernel), restated in NumPy, against the
    oracle's FilterLO.polyfilter on random subscan sets with scattered flags, heavy flags and a subscan
    that keeps only its first third."""
    import oracle

    def legvals(x, nk):
        L = np.empty((len(x), nk))
        L[:, 0] = 1.0
        if nk > 1:
            L[:, 1] = x
        for n in range(1, nk - 1):
            L[:, n + 1] = ((2 * n + 1) * x * L[:, n] - n * L[:, n - 1]) / (n + 1)
        return L

    rng = np.random.default_rng(0)
    worst, nhard = 0.0, 0
    for trial in range(80):
        order = int(rng.integers(1, 5))
        nk = order + 1
        nsub = int(rng.integers(1, 6))
        lens = rng.integers(1, 300, nsub)
        gaps = rng.integers(0, 30, nsub + 1)
        starts = np.cumsum(gaps[:-1]) + np.concatenate([[0], np.cumsum(lens[:-1])])
        nt = int(starts[-1] + lens[-1] + gaps[-1])
        npix = 20
        pix = np.sort(rng.integers(0, npix, nt)).astype(np.int64)
        mode = trial % 4
        if mode == 1:
            pix[rng.random(nt) < 0.05] = -1
        elif mode == 2:
            pix[rng.random(nt) < 0.6] = -1
        elif mode == 3:
            s = int(rng.integers(0, nsub))
            pix[starts[s] + lens[s] // 3:starts[s] + lens[s]] = -1
        phi = rng.uniform(0, np.pi, nt)
        c, sn = np.cos(2 * phi), np.sin(2 * phi)
        x = rng.normal(size=3 * npix)
        good = pix >= 0
        d = np.zeros(nt)
        d[good] = x[3 * pix[good]] + x[3 * pix[good] + 1] * c[good] + x[3 * pix[good] + 2] * sn[good]
        ref = oracle.FilterLO(nt, [lens, starts], nt, 1, pix, poly_order=order) * d
        out = np.zeros(nt)
        for a, ln in zip(starts, lens):
            m = good[a:a + ln]
            if m.sum() <= order:
                continue                                    # the reference skips the subscan
            L = legvals(np.linspace(-1, 1, ln) if ln > 1 else np.array([-1.0]), nk)
            S = np.zeros(nk)                                # moments from the run table
            t = a
            while t < a + ln:
                if pix[t] < 0:
                    t += 1
                    continue
                u = t
                while u < a + ln and pix[u] == pix[t]:
                    u += 1
                idx = np.arange(t, u)
                p = pix[t]
                S += (L[idx - a].sum(0) * x[3 * p] + (L[idx - a] * c[idx, None]).sum(0) * x[3 * p + 1]
                      + (L[idx - a] * sn[idx, None]).sum(0) * x[3 * p + 2])
                t = u
            if m.sum() == ln:
                W = np.diag(1.0 / (L * L).sum(0))
            else:
                G = L[m].T @ L[m]
                dsc = 1.0 / np.sqrt(np.diag(G))
                Gs = G * dsc[:, None] * dsc[None, :]
                try:
                    minpiv = (np.diag(np.linalg.cholesky(Gs)) ** 2).min()
                except np.linalg.LinAlgError:
                    minpiv = 0.0
                if minpiv < 0.02:                           # POLY_RUN_MIN_PIVOT: the per-subscan kernel's job
                    nhard += 1
                    out[a:a + ln] = ref[a:a + ln]
                    continue
                W = np.linalg.inv(Gs) * dsc[:, None] * dsc[None, :]
            val = d[a:a + ln] - L @ (W @ S)
            out[a:a + ln][m] = val[m]
        if good.any():
            worst = max(worst, np.abs(out - ref)[good].max() / max(np.abs(d).max(), 1e-300))
    assert worst < 1e-11, worst
    assert nhard > 0                                        # the ill-conditioned branch was exercised


# ---- round 2 host logic ---------------------------------------------------------------------------------
def test_partition_pixels_is_balanced_even_and_covers():
    from cosmomap2_b200 import distributed
    for npix in (0, 1, 7, 100, 500000, 1280001):
        for world in (1, 2, 3, 4, 8):
            lo_ = distributed.partition_pixels(npix, world)
            assert len(lo_) == world + 1 and lo_[0] == 0 and lo_[-1] == npix
            assert all(b >= a for a, b in zip(lo_[:-1], lo_[1:]))
            assert all(b % 2 == 0 for b in lo_[1:-1])          # pol * lo doubles is 16-byte aligned for any pol
            sizes = np.diff(lo_)
            assert sizes.max() - sizes.min() <= 2 + (npix % 2) or npix < 2 * world


class _FakeP(object):
    def __init__(self, nrows, ncols, pol):
        self.nrows, self.ncols, self.pol = nrows, ncols, pol


class _FakeF(object):
    def __init__(self, nsamples):
        self.nsamples = nsamples


def test_tile_order_is_chosen_from_the_map_size_and_the_timeline():
    big, small = 6000000, 500000                     # pixels: x and y together 288 MB / 24 MB at IQU
    assert lo._tod_streams(_FakeP(64 * 1000, big, 3), 1000) == 64
    assert lo._tod_streams(_FakeP(64 * 1000, small, 3), 1000) == 1          # the map stays in L2 in any order
    assert lo._tod_streams(_FakeP(64 * 1000 + 5, big, 3), 1000) == 1        # not a whole number of timelines
    assert lo._tod_streams(_FakeP(64 * 1000, big, 3), 0) == 1
    assert lo._filter_timeline(_FakeF(1000)) == 1000
    assert lo._filter_timeline(_FakeF([1000, 1000])) == 1000
    assert lo._filter_timeline(_FakeF([1000, 900])) == 0                    # CES of different lengths: time order


class _FakeZ(object):
    def __init__(self, zt):
        self._zt = zt
        self.ncols, self.nrows = zt.shape


@pytest.mark.parametrize("pol", [1, 3])
def test_banded_coarse_space_detection_is_exact(pol):
    """linearoperators._banded_coarse_space on CPU tensors: an indicator Z with A Z inside the three cyclic bands is
    compressed without loss; one entry outside, a non-indicator Z or a pixel in two columns are refused."""
    import torch
    rng = np.random.default_rng(5)
    r, npix = 6, 40
    n = pol * npix
    band = rng.integers(0, r, npix)
    band[3] = -1                                       # a pixel in no column
    Zt = torch.zeros((r, n), dtype=torch.float64)
    for p_, b in enumerate(band):
        if b >= 0:
            Zt[b, pol * p_] = 1.0
    AZt = torch.zeros((r, n), dtype=torch.float64)
    for p_, b in enumerate(band):
        if b < 0:
            continue
        for o in (-1, 0, 1):
            for k in range(pol):
                AZt[(b + o) % r, pol * p_ + k] = rng.standard_normal()
    out = lo._banded_coarse_space(_FakeZ(Zt), _FakeZ(AZt), pol)
    assert out is not None
    bd, azb = out
    assert np.array_equal(bd.numpy(), band)
    for p_, b in enumerate(band):
        for k in range(pol):
            for o in range(3):
                want = AZt[(b + o - 1) % r, pol * p_ + k].item() if b >= 0 else 0.0
                assert azb[p_, k, o].item() == want
    bad = AZt.clone()
    p0 = int(np.nonzero(band == 0)[0][0])
    bad[3, pol * p0] = 1.0                             # band 0 reaches column 3: outside its neighbours
    assert lo._banded_coarse_space(_FakeZ(Zt), _FakeZ(bad), pol) is None
    Z2 = Zt.clone()
    Z2[1, pol * p0] = 1.0                              # a pixel in two columns
    assert lo._banded_coarse_space(_FakeZ(Z2), _FakeZ(AZt), pol) is None
    Z3 = Zt.clone()
    Z3[0, pol * p0] = 0.5                              # not an indicator
    assert lo._banded_coarse_space(_FakeZ(Z3), _FakeZ(AZt), pol) is None
    if pol == 3:
        Z4 = Zt.clone()
        Z4[0, 1] = 1.0                                 # a polarisation entry
        assert lo._banded_coarse_space(_FakeZ(Z4), _FakeZ(AZt), pol) is None


@pytest.mark.parametrize("pol", [1, 3])
def test_coarse_products_probing_algebra_on_a_toy_operator(pol, monkeypatch):
    """The colouring / attribution / checksum logic of deflationlib.coarse_products is index work on tensors; here it
    runs on CPU tensors against a toy banded operator (the CUDA requirement is lifted for this test only -- no kernel
    of the library is involved).  Nearest-band coupling: 4 probes + 1 checksum apply; a coupling that skips the
    checked bands (distance 3 only) must be caught by the checksum and fall back to one apply per column."""
    import torch
    from cosmomap2_b200 import _device as dv, linop as lp, deflationlib as dl
    monkeypatch.setattr(dv, "require_cuda", lambda: None)
    r, w = 12, 40
    n = pol * r * w
    Zt = torch.zeros((r, n), dtype=torch.float64)
    for k in range(r):
        Zt[k, pol * k * w:pol * (k + 1) * w:pol] = 1.0
    dg = 1.0 + torch.rand(n, dtype=torch.float64, generator=torch.Generator().manual_seed(3))

    def coupled(shift):
        def mv(x):
            return dg * x + 0.25 * (torch.roll(x, pol * shift * w) + torch.roll(x, -pol * shift * w))
        return lp.LinearOperator(n, n, matvec=mv, symmetric=True, device=True)

    for shift, applies in ((1, 5), (3, 5 + r)):
        A = coupled(shift)
        n0 = A.nMatvec
        AZ = dl.coarse_products(A, Zt, pol)
        assert A.nMatvec - n0 == applies
        for k in range(r):
            assert torch.equal(AZ[k], A._apply(Zt[k]))
    # not an indicator space (a second non-zero per pixel): column by column from the start
    Zb = Zt.clone()
    Zb[1, 0] = 1.0
    A = coupled(1)
    n0 = A.nMatvec
    dl.coarse_products(A, Zb, pol)
    assert A.nMatvec - n0 == r
