"""
The reference's thin utility layer (utilities/linear_algebra_funcs.py,
utilities/utilities_functions.py), kept by name so its scripts and tests run unchanged.
Vector reductions and the tall-skinny products run on the device; synthetic-input generators are
plain NumPy (they only produce inputs).
"""
import random as rd
import warnings

import numpy as np
import torch

from . import _device as dv


def dgemm(A, B):
    """``dgemm(A, B) = A^T . B^T`` (utilities/linear_algebra_funcs.py:16-29), on the fp64 tensor cores
    by the library's own one-pass kernel (cm2_dense_gram): the reference calls it as
    ``dgemm(Z, Az.T)`` = ``Z^T (A Z)`` with tall-skinny Z (linearoperators.py:1019)."""
    if type(A) == list:
        A = np.asarray(A, order="F")
    if type(B) == list:
        B = np.asarray(B, order="F")
    host = not (isinstance(A, torch.Tensor) or isinstance(B, torch.Tensor))
    At = dv.to_dev_f64(np.asarray(A) if not isinstance(A, torch.Tensor) else A)
    Bt = dv.to_dev_f64(np.asarray(B) if not isinstance(B, torch.Tensor) else B)
    from . import dense
    # A: k x m, B: n x k  ->  A^T B^T (m x n) = X^T Y with X = A (k x m), Y = B^T (k x n): the kernel wants
    # the columns of X and Y contiguous, i.e. A^T (m, k) and B (n, k) row-contiguous
    out = dense.gram(At.t().contiguous(), Bt.contiguous())
    return dv.to_host(out) if host else out


def scalprod(a, b):
    """Scalar product (utilities/linear_algebra_funcs.py:39-44), deterministic device reduction."""
    host = not (isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor))
    ad = dv.to_dev_f64(np.ascontiguousarray(a) if not isinstance(a, torch.Tensor) else a)
    bd = dv.to_dev_f64(np.ascontiguousarray(b) if not isinstance(b, torch.Tensor) else b)
    out = dv.zeros_f64(1)
    dv.call("cm2_dot", dv.ptr(ad), dv.ptr(bd), ad.numel(), dv.ptr(out), dv.stream())
    return float(out.item()) if host else out[0]


def norm2(q):
    """Euclidean norm (utilities/linear_algebra_funcs.py:31-37)."""
    v = scalprod(q, q)
    return float(np.sqrt(v)) if not isinstance(v, torch.Tensor) else torch.sqrt(v)


def get_legendre_polynomials(polyorder, size):
    """utilities/linear_algebra_funcs.py:47-59 (host; only used by the out-of-scope Legendre filter)."""
    from scipy.special import legendre
    legendres = np.empty([size, polyorder + 1])
    x = np.linspace(-1, 1, size)
    for i in range(polyorder + 1):
        L = legendre(i)
        legendres[:, i] = L(x) / np.linalg.norm(L(x))
    return legendres


# ---- utilities/utilities_functions.py ---------------------------------------------------------
def is_sorted(seq):
    seq = np.asarray(seq)
    return bool(np.all(seq[:-1] <= seq[1:]))


class bash_colors:
    """utilities/utilities_functions.py:26-53."""
    HEADER = '\033[95m'
    OKBLUE = '\033[94m'
    OKGREEN = '\033[92m'
    WARNING = '\033[93m'
    FAIL = '\033[91m'
    ENDC = '\033[0m'
    BOLD = '\033[1m'
    UNDERLINE = '\033[4m'

    def header(self, string):
        return self.HEADER + str(string) + self.ENDC

    def blue(self, string):
        return self.OKBLUE + str(string) + self.ENDC

    def green(self, string):
        return self.OKGREEN + str(string) + self.ENDC

    def warning(self, string):
        return self.WARNING + str(string) + self.ENDC

    def fail(self, string):
        return self.FAIL + str(string) + self.ENDC

    def bold(self, string):
        return self.BOLD + str(string) + self.ENDC

    def underline(self, string):
        return self.UNDERLINE + str(string) + self.ENDC


def filter_warnings(wfilter):
    warnings.simplefilter(wfilter)


def angles_gen(theta0, n, sample_freq=200., whwp_freq=2.5):
    """HWP ramp theta0 + 2 pi f_hwp/f_samp i (utilities/utilities_functions.py:99-107)."""
    return theta0 + 2 * np.pi * whwp_freq / sample_freq * np.arange(n, dtype=np.float64)


def pairs_gen(nrows, ncols, rng=None):
    """utilities/utilities_functions.py:111-122."""
    if ncols < 3:
        raise RuntimeError("Not enough pixels!\n Please set Npix >=3, you have set Npix=%d" % ncols)
    if rng is not None:
        return rng.integers(0, ncols, size=nrows)
    return np.random.randint(0, high=ncols, size=nrows)


def checking_output(info):
    """utilities/utilities_functions.py:125-140."""
    if info == 0:
        return True
    if info < 0:
        raise RuntimeError("illegal input or breakdown during the execution")
    raise RuntimeError("convergence not achieved after %d iterations" % info)


def noise_val(nb, bandwidth=1, rng=None):
    """utilities/utilities_functions.py:148-177."""
    gen = np.random if rng is None else rng
    t = [gen.random(size=bandwidth) for _ in range(nb)]
    diag = [i[0] for i in t]
    return t, diag


def subscan_resize(data, subscan):
    """utilities/utilities_functions.py:179-188."""
    tmp = []
    for i in range(len(subscan[0])):
        start = subscan[1][i]
        end = subscan[1][i] + subscan[0][i]
        tmp.append(data[start:end])
    return np.concatenate(tmp)


def system_setup(nt, npix, nb, rng=None):
    """utilities/utilities_functions.py:190-212 (seeded when ``rng`` is given)."""
    gen = np.random if rng is None else rng
    d = gen.random(nt)
    pairs = pairs_gen(nt, npix, rng)
    theta0 = rd.uniform(0, np.pi) if rng is None else float(rng.uniform(0, np.pi))
    phi = angles_gen(theta0, nt)
    t, diag = noise_val(nb, 2, rng)
    return d, pairs, phi, t, diag


def reorganize_map(mapin, obspix, npix, nside, pol, fname=None):
    """utilities/healpy_functions.py:50-105 -- from the interleaved solution over the observed pixels
    to ``pol`` full-sky HEALPix arrays (12 nside^2 values each, zero where unobserved).  NumPy in ->
    list of NumPy arrays (as the reference); CUDA tensor in -> list of CUDA tensors.  Writing FITS
    files (``fname``) needs healpy, which is not part of this package."""
    if pol not in (1, 2, 3):
        raise RuntimeError("No valid polarization key set!\t=>\tpol=%d" % pol)
    if fname is not None:
        raise NotImplementedError("writing HEALPix FITS files needs healpy; pass fname=None")
    dv.require_cuda()
    hnpix = 12 * int(nside) * int(nside)
    on_dev = isinstance(mapin, torch.Tensor)
    m = dv.to_dev_f64(mapin)
    npix = int(npix)
    if m.numel() != pol * npix or len(obspix) != npix:
        raise ValueError("mapin must hold pol*npix values and obspix npix pixels")
    obs = dv.to_dev(obspix, torch.int64)
    if npix and (int(obs.min().item()) < 0 or int(obs.max().item()) >= hnpix):
        raise IndexError("obspix outside the nside=%d HEALPix range" % nside)
    out = torch.empty(pol * hnpix, dtype=torch.float64, device=m.device)
    dv.call("cm2_reorganize_map", dv.ptr(m), dv.ptr(obs), npix, pol, hnpix, dv.ptr(out),
            torch.cuda.current_stream().cuda_stream)
    parts = [out[k * hnpix:(k + 1) * hnpix] for k in range(pol)]
    return parts if on_dev else [dv.to_host(p) for p in parts]


def profile_run():
    """utilities/utilities_functions.py:65-71 -- a ``cProfile.Profile`` for the host side of a run (the
    device side is profiled with CUDA events / ncu, see tools/ and profiles/)."""
    import cProfile
    return cProfile.Profile()


def output_profile(pr):
    """utilities/utilities_functions.py:73-89 -- print the statistics collected by ``profile_run``,
    sorted by cumulative time."""
    import io
    import pstats
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats()
    print(s.getvalue())


def subtract_offset(mapp, obspix, pol):
    """utilities/healpy_functions.py:146-158 -- remove, in place, the average over the observed pixels
    (what the reference does before comparing maps solved with the offset filter, whose A has the
    monopole in its null space)."""
    if pol == 1:
        mapp[obspix] -= np.mean(mapp[obspix])
    else:
        for i in range(len(mapp)):
            mapp[i][obspix] -= np.mean(mapp[i][obspix])
    return mapp


def rescalepixels(pixs):
    """utilities/utilities_functions.py:91-96 -- ``(minpix, pixs - minpix, maxpix)``."""
    pixs = np.asarray(pixs)
    minpix, maxpix = pixs.min(), pixs.max()
    return minpix, pixs - minpix, maxpix
