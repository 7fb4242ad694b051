"""
ORACLE (test infrastructure, not product code): CPU restatement of COSMOMAP2's map-making
hot path in NumPy/SciPy, class for class and attribute for attribute.

Each class cites the reference lines it follows (paths relative to /root/reference).  The
reference itself is Python 2 + ``weave`` + ``linop`` and cannot be imported as is; the
restatement is pinned against outputs of the reference's own code run in this container through
the loader in ``oracle/refrun.py`` (fixtures under tests/golden/, generator
tests/golden/make_golden.py) and against the algebraic identities of the reference's test-suite
(SURVEY.md section 4).  Parts that delegate to the absent third-party ``krypy`` are restated
from its published algorithm in ``oracle/krylov.py`` and are *parity unpinned*.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (cosmomap2_b200/) never does.
"""
import numpy as np
from scipy.linalg import eigh, lu, solve
from scipy.linalg import get_blas_funcs

from . import linop_min as lp
from . import cloops

# --------------------------------------------------------------------------------------------
# utilities/linear_algebra_funcs.py
# --------------------------------------------------------------------------------------------


def dgemm(A, B):
    """utilities/linear_algebra_funcs.py:16-29 -- BLAS gemm(a=A.T, b=B, trans_b) == A.T @ B.T."""
    if type(A) == list:
        A = np.asarray(A, order="F")
    if type(B) == list:
        B = np.asarray(B, order="F")
    matdot = get_blas_funcs("gemm", (A, B))
    return matdot(alpha=1.0, a=A.T, b=B, trans_b=True, trans_a=False)


def get_legendre_polynomials(polyorder, size):
    """utilities/linear_algebra_funcs.py:47-59 -- columns L_k(x)/||L_k(x)||, x = linspace(-1, 1, size)."""
    from scipy.special import legendre
    legendres = np.empty([size, polyorder + 1])
    x = np.linspace(-1, 1, size)
    for i in range(polyorder + 1):
        L = legendre(i)
        legendres[:, i] = L(x) / norm2(L(x))
    return legendres


def norm2(q):
    """utilities/linear_algebra_funcs.py:31-37."""
    q = np.asarray(q)
    return get_blas_funcs("nrm2", dtype=q.dtype)(q)


def scalprod(a, b):
    """utilities/linear_algebra_funcs.py:39-44."""
    return get_blas_funcs("dot", (a, b))(a, b)


# --------------------------------------------------------------------------------------------
# utilities/utilities_functions.py (seedable ports of the synthetic-input generators)
# --------------------------------------------------------------------------------------------


def angles_gen(theta0, n, sample_freq=200., whwp_freq=2.5):
    """utilities/utilities_functions.py:99-107 -- HWP ramp theta0 + 2 pi f_hwp/f_samp * i."""
    return theta0 + 2 * np.pi * whwp_freq / sample_freq * np.arange(n, dtype=np.float64)


def pairs_gen(nrows, ncols, rng=None):
    """utilities/utilities_functions.py:111-122."""
    if ncols < 3:
        raise RuntimeError("Not enough pixels!\n Please set Npix >=3, you have set Npix=%d" % ncols)
    rng = np.random if rng is None else rng
    if hasattr(rng, "integers"):
        return rng.integers(0, ncols, size=nrows)
    return rng.randint(0, high=ncols, size=nrows)


def noise_val(nb, bandwidth=1, rng=None):
    """utilities/utilities_functions.py:148-177."""
    rng = np.random if rng is None else rng
    t = [rng.random(size=bandwidth) for _ in range(nb)]
    diag = [i[0] for i in t]
    return t, diag


def system_setup(nt, npix, nb, rng=None):
    """utilities/utilities_functions.py:190-212 (seeded when ``rng`` is a numpy Generator)."""
    rng = np.random if rng is None else rng
    d = rng.random(nt)
    pairs = pairs_gen(nt, npix, rng)
    theta0 = rng.uniform(0, np.pi)
    phi = angles_gen(theta0, nt)
    t, diag = noise_val(nb, 2, rng)
    return d, pairs, phi, t, diag


def checking_output(info):
    """utilities/utilities_functions.py:125-140."""
    if info == 0:
        return True
    if info < 0:
        raise RuntimeError("illegal input or breakdown during the execution")
    raise RuntimeError("convergence not achieved after %d iterations" % info)


def is_sorted(seq):
    seq = np.asarray(seq)
    return bool(np.all(seq[:-1] <= seq[1:]))


# --------------------------------------------------------------------------------------------
# utilities/process_ces.py
# --------------------------------------------------------------------------------------------


class ProcessTimeSamples(object):
    """utilities/process_ces.py:20-555.

    Builds the per-pixel moments of ``P^T diag(w) P``, masks unobserved / ill-conditioned pixels,
    compacts the pixel index space and relabels ``pixs`` IN PLACE (process_ces.py:416).
    """

    def __init__(self, pixs, npix, obspix=None, pol=1, phi=None, w=None, ground=None,
                 threshold_cond=1.e3, obspix2=None):
        self.pixs = pixs
        self.oldnpix = npix
        self.nsamples = len(pixs)
        self.pol = pol
        if w is None:
            w = np.ones(self.nsamples)
        if obspix is None:
            obspix = np.arange(self.nsamples)          # process_ces.py:67-68 (sic: nsamples)
        self.obspix = obspix
        if ground is not None:                          # :70-73
            neg = ground < 0
            ground[neg] = -1
            pixs[neg] = -1
        if obspix2 is None:
            self.threshold = threshold_cond
            self.initializeweights(phi, w)
            self.new_repixelization()
            self.flagging_samples()
        else:
            self.SetObspix(obspix2)
            self.flagging_samples()
            self.compute_arrays(phi, w)
        if ground is not None:                          # :85-89
            ground[pixs == -1] = -1
            self.ground = ground

    @property
    def get_new_pixel(self):
        return self.__new_npix, self.obspix

    # -- :94-111
    def SetObspix(self, new_obspix):
        self.old2new = np.full(self.oldnpix, -1, dtype=np.int32)
        if not (is_sorted(self.obspix) and is_sorted(new_obspix)):
            indexsorted = np.argsort(self.obspix, kind="quicksort")
            self.obspix = self.obspix[indexsorted]
        idx = np.searchsorted(self.obspix, new_obspix)
        self.old2new[idx] = np.arange(len(idx))
        self.obspix = new_obspix
        self.__new_npix = len(new_obspix)

    def _accumulate(self, npix, phi, w):
        """The weave loops #10-15 (process_ces.py:125-189, 480-542): per-pixel weighted moments.
        ``np.bincount`` adds in sample order, i.e. in the same order as the serial C loops."""
        pixs = np.asarray(self.pixs)
        good = pixs != -1
        p = pixs[good]
        wg = np.asarray(w, dtype=np.float64)[good]
        if self.pol == 1:
            self.counts = np.bincount(p, weights=wg, minlength=npix).astype(np.float64)
            return
        self.cos = np.cos(2. * phi)
        self.sin = np.sin(2. * phi)
        c, s = self.cos[good], self.sin[good]
        self.cos2 = np.bincount(p, weights=wg * c * c, minlength=npix).astype(np.float64)
        self.sin2 = np.bincount(p, weights=wg * s * s, minlength=npix).astype(np.float64)
        self.sincos = np.bincount(p, weights=wg * s * c, minlength=npix).astype(np.float64)
        if self.pol == 3:
            self.counts = np.bincount(p, weights=wg, minlength=npix).astype(np.float64)
            self.cosine = np.bincount(p, weights=wg * c, minlength=npix).astype(np.float64)
            self.sine = np.bincount(p, weights=wg * s, minlength=npix).astype(np.float64)

    # -- :113-189
    def compute_arrays(self, phi, w):
        self._accumulate(self.__new_npix, phi, w)

    # -- :426-555
    def initializeweights(self, phi, w):
        self._accumulate(self.oldnpix, phi, w)
        if self.pol == 1:
            self.mask = np.where(self.counts > 0)[0]
            return
        with np.errstate(all="ignore"):
            det = (self.cos2 * self.sin2) - (self.sincos * self.sincos)
            tr = self.cos2 + self.sin2
            sqrt = np.sqrt(tr * tr / 4. - det)
            lambda_max = tr / 2. + sqrt
            lambda_min = tr / 2. - sqrt
            cond_num = np.abs(lambda_max / lambda_min)
            mask = np.where(cond_num <= self.threshold)[0]
        if self.pol == 2:
            self.mask = mask
        else:
            mask2 = np.where(self.counts > 2)[0]
            self.mask = np.intersect1d(mask2, mask)

    # -- :192-349.  The reference does an O(Nold*Nmask) membership search per old pixel; the
    # result is "keep the masked-in pixels in ascending order", restated with a boolean mask.
    def new_repixelization(self):
        keep = np.zeros(self.oldnpix, dtype=bool)
        keep[self.mask] = True
        n_new = int(keep.sum())
        old2new = np.full(self.oldnpix, -1, dtype=int)
        old2new[keep] = np.arange(n_new)
        names = {1: ["counts"], 2: ["cos2", "sin2", "sincos"],
                 3: ["cos2", "sin2", "sincos", "cosine", "sine", "counts"]}[self.pol]
        for nm in names:
            setattr(self, nm, getattr(self, nm)[keep].copy())
        # obspix(Nnew)=obspix(jpix) for kept jpix, then np.delete(obspix, range(n_new, oldnpix))
        obspix = np.asarray(self.obspix)
        head = obspix[:self.oldnpix][keep[:len(obspix[:self.oldnpix])]]
        self.obspix = np.concatenate([head, obspix[self.oldnpix:]])
        self.old2new = old2new
        self.__new_npix = n_new

    # -- :403-425
    def flagging_samples(self):
        pixs = self.pixs
        good = pixs != -1
        pixs[good] = self.old2new[pixs[good]]


# --------------------------------------------------------------------------------------------
# interfaces/linearoperators.py
# --------------------------------------------------------------------------------------------


USE_C_LOOPS = True   # route the per-sample loops through oracle/weave_loops.c when it is built


def _c():
    return USE_C_LOOPS and cloops.available()


class SparseLO(lp.LinearOperator):
    """interfaces/linearoperators.py:326-557 -- pointing operator P and its transpose.
    Two equivalent restatements: NumPy (bincount adds in sample order) and the plain-C twin of
    the weave loops (oracle/weave_loops.c); tests/test_oracle_golden.py checks they agree bit for
    bit.  The C twin is what bench.py times as the CPU baseline."""

    def _cmult(self, v):
        return cloops.pointing_mult(self.pairs, getattr(self, "cos", None), getattr(self, "sin", None),
                                    self.pol, np.ascontiguousarray(v, dtype=np.float64))

    def _crmult(self, v):
        return cloops.pointing_rmult(self.pairs, getattr(self, "cos", None), getattr(self, "sin", None),
                                     self.pol, np.ascontiguousarray(v, dtype=np.float64), self.ncols)

    def mult(self, v):                                   # :356-384
        if _c():
            return self._cmult(v)
        x = np.zeros(self.nrows)
        g = self.pairs != -1
        x[g] = v[self.pairs[g]]
        return x

    def rmult(self, v):                                  # :385-410
        if _c():
            return self._crmult(v)
        g = self.pairs != -1
        return np.bincount(self.pairs[g], weights=v[g], minlength=self.ncols).astype(np.float64)

    def mult_qu(self, v):                                # :411-438
        if _c():
            return self._cmult(v)
        x = np.zeros(self.nrows)
        g = self.pairs != -1
        p = self.pairs[g]
        x[g] = v[2 * p] * self.cos[g] + v[2 * p + 1] * self.sin[g]
        return x

    def rmult_qu(self, v):                               # :439-462
        if _c():
            return self._crmult(v)
        out = np.zeros(self.ncols * self.pol)
        g = self.pairs != -1
        p = self.pairs[g]
        out[0::2] = np.bincount(p, weights=v[g] * self.cos[g], minlength=self.ncols)
        out[1::2] = np.bincount(p, weights=v[g] * self.sin[g], minlength=self.ncols)
        return out

    def mult_iqu(self, v):                               # :463-497
        if _c():
            return self._cmult(v)
        x = np.zeros(self.nrows)
        g = self.pairs != -1
        p = self.pairs[g]
        x[g] = v[3 * p] + v[3 * p + 1] * self.cos[g] + v[3 * p + 2] * self.sin[g]
        return x

    def rmult_iqu(self, v):                              # :498-526
        if _c():
            return self._crmult(v)
        out = np.zeros(self.ncols * self.pol)
        g = self.pairs != -1
        p = self.pairs[g]
        out[0::3] = np.bincount(p, weights=v[g], minlength=self.ncols)
        out[1::3] = np.bincount(p, weights=v[g] * self.cos[g], minlength=self.ncols)
        out[2::3] = np.bincount(p, weights=v[g] * self.sin[g], minlength=self.ncols)
        return out

    def __init__(self, n, m, pix_samples, pol=1, angle_processed=None):   # :527-550
        self.ncols = n
        self.nrows = m
        self.pol = pol
        self.pairs = pix_samples
        if self.pol > 1:
            self.cos = angle_processed.cos
            self.sin = angle_processed.sin
        if pol == 3:
            self.__runcase = "IQU"
            super(SparseLO, self).__init__(nargin=self.pol * self.ncols, nargout=self.nrows,
                                           matvec=self.mult_iqu, symmetric=False,
                                           rmatvec=self.rmult_iqu)
        elif pol == 1:
            self.__runcase = "I"
            super(SparseLO, self).__init__(nargin=self.pol * self.ncols, nargout=self.nrows,
                                           matvec=self.mult, symmetric=False, rmatvec=self.rmult)
        elif pol == 2:
            self.__runcase = "QU"
            super(SparseLO, self).__init__(nargin=self.pol * self.ncols, nargout=self.nrows,
                                           matvec=self.mult_qu, symmetric=False,
                                           rmatvec=self.rmult_qu)
        else:
            raise RuntimeError("No valid polarization key set!\t=>\tpol=%d \n \
                                    Possible values are pol=%d(I),%d(QU), %d(IQU)." % (pol, 1, 2, 3))

    @property
    def maptype(self):
        return self.__runcase


class ToeplitzLO(lp.LinearOperator):
    """interfaces/linearoperators.py:560-602 -- symmetric banded Toeplitz, zero boundaries."""

    def mult(self, v):                                   # :582-595
        if cloops.available() and len(self.array) > 8:
            return cloops.toeplitz(np.ascontiguousarray(self.array, dtype=np.float64),
                                   np.ascontiguousarray(v, dtype=np.float64))
        y = self.array[0] * v
        for i in range(1, len(self.array)):
            if i >= len(v):
                break
            temp = self.array[i] * v
            y[:-i] += temp[i:]
            y[i:] += temp[:-i]
        return y

    def __init__(self, a, size):                         # :598-602
        super(ToeplitzLO, self).__init__(nargin=size, nargout=size, matvec=self.mult,
                                         symmetric=True)
        self.array = a


class WeightingLO(lp.LinearOperator):
    """interfaces/linearoperators.py:604-625 -- per-(CES, detector) scalar weight; IN PLACE."""

    def mult(self, d):
        offset = 0
        oldb = 0
        for b, ns in zip(self.ndet_pairs, self.nsample_per_pair):
            for idx, w in np.ndenumerate(self.weights[oldb:b + oldb]):
                istart = idx[0] * ns + offset
                iend = (idx[0] + 1) * ns + offset
                d[istart:iend] = w * d[istart:iend]
            offset += b * ns
            oldb += b
        return d

    def __init__(self, bolos_per_ces, samples_per_bolopair, weights):
        self.ndet_pairs = bolos_per_ces
        self.nsample_per_pair = samples_per_bolopair
        self.size = int(np.sum([i * j for i, j in zip(samples_per_bolopair, bolos_per_ces)]))
        self.weights = np.asarray(weights)
        super(WeightingLO, self).__init__(nargin=self.size, nargout=self.size,
                                          matvec=self.mult, symmetric=True)


class BlockDiagonalLinearOperator(lp.LinearOperator):
    """interfaces/blkop.py:140-242 -- block diagonal of operators, applied block by block."""

    def __init__(self, blocks, **kwargs):
        try:
            for block in blocks:
                block.shape
        except (TypeError, AttributeError):
            raise ValueError("blocks should be a flattened list of operators")
        symmetric = all(b.symmetric for b in blocks)
        self._blocks = blocks
        nargin = sum(b.shape[-1] for b in blocks)
        nargout = sum(b.shape[0] for b in blocks)
        blocksT = [b.T for b in blocks]

        def blk_matvec(x, blks):
            nargins = [b.shape[-1] for b in blocks]
            nargouts = [b.shape[0] for b in blocks]
            if len(x) != sum(nargins):
                raise lp.ShapeError("Multiplying with vector of wrong shape.")
            y = np.empty(sum(nargouts), dtype=np.result_type(self.dtype, x.dtype))
            r0 = c0 = 0
            for k, B in enumerate(blks):
                r1 = r0 + nargouts[k]
                c1 = c0 + nargins[k]
                y[r0:r1] = B * x[c0:c1]
                r0, c0 = r1, c1
            return y

        kwargs.pop("symmetric", None)
        super(BlockDiagonalLinearOperator, self).__init__(
            nargin, nargout, symmetric=symmetric,
            matvec=lambda x: blk_matvec(x, self._blocks),
            rmatvec=lambda x: blk_matvec(x, blocksT), **kwargs)

    @property
    def blocks(self):
        return self._blocks

    def __getitem__(self, idx):
        blks = self._blocks[idx]
        if isinstance(idx, slice):
            return BlockDiagonalLinearOperator(blks)
        return blks


class BlockLO(BlockDiagonalLinearOperator):
    """interfaces/linearoperators.py:627-697 -- N^-1 as ``nblocks`` equal-size blocks.

    ``offdiag=False``: block i = t[i] * identity, ``self.diag`` = per-sample weight vector.
    ``offdiag=True``:  block i = ToeplitzLO(t[i], blocksize); ``self.diag`` = t[0] (:673, sic).
    Deviation kept from the reference's *tests*: they also pass ``blocksize`` as a list of
    per-block sizes (tests/test_toeplitz_vector_multiplication.py:12), which the reference's
    ``np.ones(self.blocksize)`` cannot honour; here a list means variable block sizes.
    """

    def build_blocks(self):
        tmplist = []
        self.blocklist = []
        nb = len(self.covnoise)
        sizes = list(self.blocksize) if np.ndim(self.blocksize) else [int(self.blocksize)] * nb
        if self.isoffdiag:
            tmplist.append(np.atleast_1d(np.asarray(self.covnoise[0], dtype=np.float64)))
            self.blocklist = [ToeplitzLO(np.atleast_1d(a), sz) for a, sz in zip(self.covnoise, sizes)]
        else:
            for val, sz in zip(self.covnoise, sizes):
                d = np.full(sz, val, dtype=np.float64)
                self.blocklist.append(lp.DiagonalOperator(d))
                tmplist.append(d)
        self.diag = np.concatenate(tmplist)

    def __init__(self, blocksize, t, offdiag=False):
        self.__isoffdiag = offdiag
        self.blocksize = blocksize
        self.covnoise = t
        self.build_blocks()
        super(BlockLO, self).__init__(self.blocklist)

    @property
    def isoffdiag(self):
        return self.__isoffdiag


class FilterLO(lp.LinearOperator):
    """interfaces/linearoperators.py:94-168, 263-282 -- subscan offset removal (poly_order=0).

    TOD layout: CES-major, detector-major, time-minor (:134-140, 167)."""

    def mult(self, d):                                   # :129-168
        vec_out = d * 0.
        pixs = self.pixels
        offset = 0
        for subsc, ts, ns, nb in zip(self.subscans, self.tstart, self.nsamples, self.nbolos):
            n = nb * ns
            for bolo_iter in range(nb):
                for i, j in zip(subsc, ts):
                    start = int(j + (ns * bolo_iter) + offset)
                    end = int(start + i)
                    seg = d[start:end]
                    good = pixs[start:end] != -1
                    cnt = np.count_nonzero(good)
                    if cnt == 0:                        # mean = 0/0 -> nan -> skipped (:163-164)
                        continue
                    dmean = _seq_sum(seg[good]) / float(cnt)
                    if np.isinf(dmean) or np.isnan(dmean):
                        continue
                    vec_out[start:end] = seg - dmean
            offset += n
        return vec_out

    def __init__(self, size, subscan_nsample, samples_per_bolopair, bolos_per_ces, pix_samples,
                 poly_order=0, npool=4):
        self.n = size
        self.nsamples = samples_per_bolopair
        self.nbolos = bolos_per_ces
        self.subscans = subscan_nsample[0]
        self.tstart = subscan_nsample[1]
        if not (type(self.nsamples) is list):
            self.nsamples = [self.nsamples]
            self.nbolos = [self.nbolos]
            self.subscans = [self.subscans]
            self.tstart = [self.tstart]
        self.pixels = pix_samples
        self.poly_order = poly_order
        if poly_order == 0:
            super(FilterLO, self).__init__(nargin=size, nargout=size, matvec=self.mult,
                                           symmetric=False)
        else:                                            # :279-282 (the Pool is an implementation detail)
            self.compute_legendres()
            super(FilterLO, self).__init__(nargin=size, nargout=size, matvec=self.polyfilter,
                                           symmetric=False)

    def compute_legendres(self):                         # :206-213
        sizes = []
        for array in self.subscans:
            for i in array:
                if int(i) not in sizes:
                    sizes.append(int(i))
        self.legendres = {size: get_legendre_polynomials(self.poly_order, size) for size in sizes}

    def polyfilter(self, d):                             # :170-204 (== procsfilter/globalprocsfilter per CES)
        vec_out = d * 0.
        mask = np.asarray(self.pixels) >= 0
        offset = 0
        for subsc, ts, ns, nb in zip(self.subscans, self.tstart, self.nsamples, self.nbolos):
            n = nb * ns
            for bolo_iter in range(nb):
                for i, j in zip(subsc, ts):
                    start = int(j + (ns * bolo_iter) + offset)
                    end = int(start + i)
                    tmpmask = mask[start:end]
                    size = int(np.count_nonzero(tmpmask))
                    if size <= self.poly_order:
                        continue
                    legendres = self.legendres[int(i)]
                    if size != i:
                        q, r = np.linalg.qr(legendres[tmpmask])
                        legendres = q
                    p = np.zeros(size)
                    seg = d[start:end][tmpmask]
                    for k in range(self.poly_order + 1):
                        filterbasis = legendres[:, k]
                        p += scalprod(np.ascontiguousarray(filterbasis), np.ascontiguousarray(seg)) * filterbasis
                    out = vec_out[start:end]
                    out[tmpmask] = seg - p
            offset += n
        return vec_out


def _seq_sum(a):
    """Left-to-right fp64 sum, as the serial C loop at linearoperators.py:147-153 does."""
    if cloops.available():
        return cloops.seq_sum(np.ascontiguousarray(a, dtype=np.float64))
    if len(a) == 0:
        return 0.0
    return float(np.cumsum(a)[-1])


class BlockDiagonalLO(lp.LinearOperator):
    """interfaces/linearoperators.py:700-746 -- explicit P^T diag(N^-1) P, per-pixel blocks."""

    def __init__(self, CES, n, pol=1):
        self.size = pol * n
        self.pol = pol
        super(BlockDiagonalLO, self).__init__(nargin=self.size, nargout=self.size,
                                              matvec=self.mult, symmetric=True)
        self.pixels = np.arange(n)
        if pol == 1:
            self.counts = CES.counts
        elif pol > 1:
            self.sin2 = CES.sin2
            self.sincos = CES.sincos
            self.cos2 = CES.cos2
            if pol == 3:
                self.counts = CES.counts
                self.cos = CES.cosine
                self.sin = CES.sine

    def mult(self, x):                                   # :728-746
        y = x * 0.
        if self.pol == 1:
            y = x * self.counts
        elif self.pol == 3:
            h, c, s, c2, s2, cs = self.counts, self.cos, self.sin, self.cos2, self.sin2, self.sincos
            y[0::3] = h * x[0::3] + c * x[1::3] + s * x[2::3]
            y[1::3] = c * x[0::3] + c2 * x[1::3] + cs * x[2::3]
            y[2::3] = s * x[0::3] + cs * x[1::3] + s2 * x[2::3]
        elif self.pol == 2:
            c2, s2, cs = self.cos2, self.sin2, self.sincos
            y[0::2] = c2 * x[0::2] + cs * x[1::2]
            y[1::2] = cs * x[0::2] + s2 * x[1::2]
        return y


class BlockDiagonalPreconditionerLO(lp.LinearOperator):
    """interfaces/linearoperators.py:749-859 -- M_BD = (P^T diag(N^-1) P)^-1, closed form."""

    def mult(self, x):                                   # :775-841
        if _c() and self.pol in (2, 3):
            n = self.size // self.pol
            g = lambda nm: np.ascontiguousarray(getattr(self, nm), dtype=np.float64) if hasattr(self, nm) else None  # noqa: E731
            return cloops.bd_apply(n, self.pol, g("counts"), g("cos"), g("sin"), g("cos2"), g("sin2"),
                                   g("sincos"), np.ascontiguousarray(x, dtype=np.float64))
        y = x * 0.
        if self.pol == 1:
            m = self.counts > 0
            y[m] = x[m] / self.counts[m]
        elif self.pol == 3:
            h, c, s, c2, s2, cs = self.counts, self.cos, self.sin, self.cos2, self.sin2, self.sincos
            det = h * (c2 * s2 - cs * cs) - c * c * s2 - s * s * c2 + 2. * c * s * cs
            m = np.abs(det) > 1e-5
            x0, x1, x2 = x[0::3], x[1::3], x[2::3]
            with np.errstate(all="ignore"):
                y0 = ((c2 * s2 - cs * cs) * x0 + (s * cs - c * s2) * x1 + (c * cs - s * c2) * x2) / det
                y1 = ((s * cs - c * s2) * x0 + (h * s2 - s * s) * x1 + (s * c - h * cs) * x2) / det
                y2 = ((c * cs - s * c2) * x0 + (-h * cs + c * s) * x1 + (h * c2 - c * c) * x2) / det
            y[0::3] = np.where(m, y0, 0.)
            y[1::3] = np.where(m, y1, 0.)
            y[2::3] = np.where(m, y2, 0.)
        elif self.pol == 2:
            c2, s2, cs = self.cos2, self.sin2, self.sincos
            det = (c2 * s2) - (cs * cs)
            m = np.abs(det) > 1e-5
            x0, x1 = x[0::2], x[1::2]
            with np.errstate(all="ignore"):
                y0 = (s2 * x0 - cs * x1) / det
                y1 = (-cs * x0 + c2 * x1) / det
            y[0::2] = np.where(m, y0, 0.)
            y[1::2] = np.where(m, y1, 0.)
        return y

    def __init__(self, CES, n, pol=1):                   # :843-859
        self.size = pol * n
        self.pixels = np.arange(n)
        self.pol = pol
        if pol == 1:
            self.counts = CES.counts
        elif pol > 1:
            self.sin2 = CES.sin2
            self.cos2 = CES.cos2
            self.sincos = CES.sincos
            if pol == 3:
                self.counts = CES.counts
                self.cos = CES.cosine
                self.sin = CES.sine
        super(BlockDiagonalPreconditionerLO, self).__init__(nargin=self.size, nargout=self.size,
                                                            matvec=self.mult, symmetric=True)


class InverseLO(lp.LinearOperator):
    """interfaces/linearoperators.py:861-941 -- a SciPy solver wrapped as A^-1 (host glue)."""

    def mult(self, x):
        y, info = self.method(self.A, x, M=self.preconditioner)
        self.isconverged(info)
        return y

    def isconverged(self, info):
        self.__converged = info
        return info == 0

    def __init__(self, A, method=None, preconditioner=None):
        super(InverseLO, self).__init__(nargin=A.shape[0], nargout=A.shape[1], matvec=self.mult,
                                        symmetric=True)
        self.A = A
        self.__method = method
        self.__preconditioner = preconditioner
        self.__converged = None

    method = property(lambda self: self.__method)
    converged = property(lambda self: self.__converged)
    preconditioner = property(lambda self: self.__preconditioner)


class CoarseLO(lp.LinearOperator):
    """interfaces/linearoperators.py:946-1027 -- E = Z^T A Z and the action of E^-1."""

    def mult(self, v):                                   # :969-977
        y = solve(self.L, v, lower=True, overwrite_b=False)
        return solve(self.U, y, overwrite_b=True)

    def mult_eig(self, v):                               # :979-984
        return self.invE.dot(v)

    def setting_inverse_w_eigenvalues(self, E):          # :986-1015
        eigenvals, W = eigh(E)
        lambda_max = max(eigenvals)
        diags = eigenvals * 0.
        threshold_to_degen = 1.e-6
        nondegenerate = np.where(abs(eigenvals / lambda_max) > threshold_to_degen)[0]
        self.ndiscarded = len(eigenvals) - len(nondegenerate)
        for i in nondegenerate:
            diags[i] = 1. / eigenvals[i]
        D = np.diag(diags)
        tmp = dgemm(D.T, W)
        self.invE = dgemm(W.T, tmp.T)

    def __init__(self, Z, Az, r, apply="LU"):            # :1018-1027
        M = dgemm(Z, Az.T)
        self.E = M.copy()
        if apply == "eig":
            self.setting_inverse_w_eigenvalues(M)
            super(CoarseLO, self).__init__(nargin=r, nargout=r, matvec=self.mult_eig,
                                           symmetric=True)
        elif apply == "LU":
            self.L, self.U = lu(M, permute_l=True, overwrite_a=True, check_finite=False)
            super(CoarseLO, self).__init__(nargin=r, nargout=r, matvec=self.mult, symmetric=True)


class DeflationLO(lp.LinearOperator):
    """interfaces/linearoperators.py:1029-1065 -- Z y and Z^T x, column by column."""

    def mult(self, x):                                   # :1041-1050
        y = np.zeros(self.nrows)
        for i in range(self.ncols):
            y += self.z[i] * x[i]
        return y

    def rmult(self, x):                                  # :1051-1056
        return np.array([scalprod(np.ascontiguousarray(i), x) for i in self.z])

    def __init__(self, z):
        self.z = []
        self.nrows, self.ncols = z.shape
        z = np.asarray(z)
        for j in range(self.ncols):
            self.z.append(z[:, j])
        super(DeflationLO, self).__init__(nargin=self.ncols, nargout=self.nrows,
                                          matvec=self.mult, symmetric=False, rmatvec=self.rmult)


class GroundFilterLO(lp.LinearOperator):
    """interfaces/linearoperators.py:24-61 -- I - G (G^T G)^-1 G^T over ground-template bins."""

    def counts_in_groundbins(self, g):                   # :26-46
        g = np.asarray(g)
        return np.bincount(g[g != -1], minlength=self.nbins).astype(np.float64)

    def mult(self, v):                                   # :48-49
        return v - self.Pg * v

    def __init__(self, ground):                          # :51-61
        self.nbins = int(max(ground)) + 1
        self.n = len(ground)
        counts = self.counts_in_groundbins(ground)
        G = SparseLO(self.nbins, self.n, ground)
        G.counts = counts
        invGtG = BlockDiagonalPreconditionerLO(G, self.nbins)
        self.Pg = (G * invGtG * G.T)
        super(GroundFilterLO, self).__init__(nargin=self.n, nargout=self.n, matvec=self.mult,
                                             symmetric=True)


def reorganize_map(mapin, obspix, npix, nside, pol, fname=None):
    """utilities/healpy_functions.py:50-105 -- de-interleave the solution and expand the observed
    pixels to full-sky HEALPix arrays (hp.nside2npix(nside) = 12 nside^2); writing to file is out
    of scope."""
    healpix_npix = 12 * nside * nside
    obspix = np.asarray(obspix)
    if pol == 3:
        m = np.zeros((healpix_npix, 3))
        m[obspix, 0], m[obspix, 1], m[obspix, 2] = mapin[::3], mapin[1::3], mapin[2::3]
        return [m[:, 0], m[:, 1], m[:, 2]]
    if pol == 2:
        m = np.zeros((healpix_npix, 2))
        m[obspix, 0], m[obspix, 1] = mapin[::2], mapin[1::2]
        return [m[:, 0], m[:, 1]]
    m = np.zeros(healpix_npix)
    m[obspix] = mapin
    return [m]


def two_level_preconditioner(Mbd, A_or_AZd, Zd, E, n):
    """The composition at src/test_M2_precond_onto_real_data.py:109-112 and
    tests/test_2level_preconditioner.py:44-48: R = I - AZd*E*Zd.T ; M2 = Mbd*R + Zd*E*Zd.T.
    ``A_or_AZd`` is either DeflationLO(A Z) or the product ``A*Zd``."""
    I = lp.IdentityOperator(n)
    R = I - A_or_AZd * E * Zd.T
    return Mbd * R + Zd * E * Zd.T, R
