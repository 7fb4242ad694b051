"""
``ProcessTimeSamples`` -- device-side TOD pre-processing with the reference's constructor,
attributes and in-place side effects (utilities/process_ces.py:20-555).

What runs where: the per-sample passes (cos/sin of 2 phi, weighted per-pixel moments, pixel
relabelling) and the per-pixel passes (good-pixel mask, prefix-sum compaction that replaces the
reference's O(Nold*Nmask) search at :192-349) are CUDA kernels behind the C ABI; only O(npix)
integer bookkeeping of ``obspix`` and the final copy of the relabelled pixels back into the
caller's array (the reference mutates ``pixs`` in place, :416) happen on the host.
"""
import numpy as np
import torch

from . import _device as dv


# Strict parity: cos/sin of 2 phi taken from NumPy on the host and the moments accumulated in the
# serial loop's order and rounding, so counts/cosine/.../sincos, the mask, old2new and the relabelled
# pixels are BIT-IDENTICAL to the reference on the same inputs (the default path sums with atomics:
# equal to ~1e-16, masks equal except on rounding knife edges).  One-off set-up cost:
# a stable sort of the samples by pixel and a host cos/sin.
STRICT_PARITY = bool(int(__import__("os").environ.get("CM2_STRICT_PARITY", "0")))


class BlockWeights(object):
    """Per-sample white-noise weights described by per-block scalars (``BlockLO.diag``).

    Behaves as the NumPy vector the reference builds at linearoperators.py:676-683 (``__array__``,
    ``len``, indexing) but lets ``ProcessTimeSamples(w=N.diag)`` stay O(nblocks) on the device.
    """

    def __init__(self, values, starts):
        self.values = np.asarray(values, dtype=np.float64)
        self.starts = np.asarray(starts, dtype=np.int64)      # nblocks + 1
        self._dense = None
        self._values_dev = None
        self._starts_dev = None

    def __len__(self):
        return int(self.starts[-1])

    @property
    def shape(self):
        return (len(self),)

    dtype = np.dtype(np.float64)
    ndim = 1

    def equal_blocksize(self):
        sizes = np.diff(self.starts)
        return int(sizes[0]) if len(sizes) and np.all(sizes == sizes[0]) else 0

    def __array__(self, dtype=None, copy=None):
        if self._dense is None:
            self._dense = np.repeat(self.values, np.diff(self.starts))
        return self._dense if dtype is None else self._dense.astype(dtype)

    def __getitem__(self, idx):
        return self.__array__()[idx]

    def copy(self):
        return self.__array__().copy()

    def device_args(self):
        """(wblk_ptr, nblocks, blocksize, blk_start_ptr) for the C ABI."""
        if self._values_dev is None:
            self._values_dev = dv.to_dev_f64(self.values)
            if not self.equal_blocksize():
                self._starts_dev = dv.to_dev(self.starts, torch.int64)
        return (dv.ptr(self._values_dev), len(self.values), self.equal_blocksize(),
                dv.ptr(self._starts_dev))


class ProcessTimeSamples(object):
    """Same signature and attributes as the reference class (utilities/process_ces.py:58).

    Attributes kept: ``counts, cosine, sine, cos2, sin2, sincos`` (per good pixel), ``cos, sin``
    (per sample), ``mask, old2new, obspix, pixs, oldnpix, nsamples, pol`` and the property
    ``get_new_pixel -> (npix_new, obspix)``.  ``pixs`` (and ``ground``) are modified IN PLACE.
    Host views of device arrays are materialised lazily on first access.
    """

    def __init__(self, pixs, npix, obspix=None, pol=1, phi=None, w=None, ground=None,
                 threshold_cond=1.e3, obspix2=None, comm=None):
        """``comm`` (extension, default None = the reference's single-process behaviour): a
        torch.distributed process group, or True for the default group.  The per-pixel moments are
        then summed over ranks before the good-pixel mask is taken, so every rank of a
        detector-sharded run derives the same pixel set."""
        dv.require_cuda()
        self._comm = comm
        if pol not in (1, 2, 3):
            raise RuntimeError("No valid polarization key set!\t=>\tpol=%d \n "
                               "Possible values are pol=%d(I),%d(QU), %d(IQU)." % (pol, 1, 2, 3))
        self.pixs = pixs
        self.oldnpix = int(npix)
        self.nsamples = len(pixs)
        self.pol = pol
        self._host = {}
        if obspix is None:
            obspix = np.arange(self.nsamples)                  # process_ces.py:67-68 (sic)
        self.obspix = obspix
        if ground is not None:                                  # :70-73 (host: touches caller arrays)
            neg = np.asarray(ground) < 0
            ground[neg] = -1
            pixs[neg] = -1
        st = dv.stream()
        nt = self.nsamples
        self._pix_dev = dv.pix_to_dev(pixs)
        if isinstance(pixs, np.ndarray) and dv.lookup_host_image(pixs) is self._pix_dev:
            self._pix_dev = self._pix_dev.clone()               # never relabel another object's image
        # angles
        self._cos_dev = self._sin_dev = None
        if pol > 1:
            if phi is None:
                raise ValueError("phi is required for pol=2,3")
            if STRICT_PARITY:
                phi_h = dv.to_host(phi) if isinstance(phi, torch.Tensor) else np.asarray(phi, dtype=np.float64)
                self._host["cos"], self._host["sin"] = np.cos(2. * phi_h), np.sin(2. * phi_h)   # :493-494
                self._cos_dev, self._sin_dev = dv.to_dev_f64(self._host["cos"]), dv.to_dev_f64(self._host["sin"])
            else:
                phi_dev = dv.to_dev_f64(phi)
                self._cos_dev = dv.empty_f64(nt)
                self._sin_dev = dv.empty_f64(nt)
                dv.call("cm2_angles", dv.ptr(phi_dev), nt, dv.ptr(self._cos_dev), dv.ptr(self._sin_dev), st)
                del phi_dev
        self._w = w
        if obspix2 is None:
            self.threshold = threshold_cond
            mom = self._moments(self.oldnpix)                  # initializeweights :426-555
            self._reduce(mom)
            good = torch.empty(self.oldnpix, dtype=torch.int32, device=mom.device)
            dv.call("cm2_weights_mask", dv.ptr(mom), self.oldnpix, pol, float(threshold_cond), dv.ptr(good), st)
            self._repixelize(mom, good)                        # new_repixelization :192-349
            self._relabel()                                    # flagging_samples :403-425
        else:
            self._set_obspix(obspix2)                          # SetObspix :94-111
            self._relabel()
            self._mom_dev = self._moments(self.__new_npix)     # compute_arrays :113-189
            self._reduce(self._mom_dev)
        self._write_back()
        if ground is not None:                                  # :85-89
            ground[np.asarray(pixs) == -1] = -1
            self.ground = ground

    # ---- device passes ---------------------------------------------------------------------
    def _moments(self, npix):
        st = dv.stream()
        mom = torch.empty(max(npix, 1) * 6, dtype=torch.float64, device=self._pix_dev.device)
        w = self._w
        wptr, blk = None, (None, 0, 0, None)
        keep = None
        if isinstance(w, BlockWeights):
            blk = w.device_args()
        elif w is not None:
            keep = dv.to_dev_f64(w)
            if keep.numel() != self.nsamples:
                raise ValueError("w must have one weight per sample")
            wptr = dv.ptr(keep)
        if STRICT_PARITY:
            pix = self._pix_dev
            order = torch.sort(pix, stable=True).indices
            nflag = int((pix < 0).sum().item())
            perm = order[nflag:].to(torch.int32).contiguous()
            rowptr = torch.zeros(npix + 1, dtype=torch.int64, device=pix.device)
            rowptr[1:] = torch.cumsum(torch.bincount(pix[pix >= 0].to(torch.int64), minlength=npix), 0)
            dv.call("cm2_weights_moments_sorted", dv.ptr(rowptr), dv.ptr(perm), dv.ptr(self._cos_dev),
                    dv.ptr(self._sin_dev), wptr, blk[0], blk[1], blk[2], blk[3], self.pol, dv.ptr(mom), npix, st)
            return mom
        dv.call("cm2_weights_moments", dv.ptr(self._pix_dev), dv.ptr(self._cos_dev), dv.ptr(self._sin_dev),
                wptr, blk[0], blk[1], blk[2], blk[3], self.nsamples, self.pol, dv.ptr(mom), npix, st)
        return mom

    def _reduce(self, mom):
        if self._comm is not None and self._comm is not False:
            from . import distributed
            distributed.all_reduce_sum_(mom, None if self._comm is True else self._comm)

    def _repixelize(self, mom, good):
        st = dv.stream()
        n = self.oldnpix
        old2new = torch.empty(max(n, 1), dtype=torch.int32, device=mom.device)
        count = torch.zeros(1, dtype=torch.int64, device=mom.device)
        scratch = torch.empty(int(dv.call("cm2_scan_scratch_bytes", n)) // 8 + 1, dtype=torch.int64,
                              device=mom.device)
        dv.call("cm2_weights_old2new", dv.ptr(good), n, dv.ptr(old2new), dv.ptr(count), dv.ptr(scratch), st)
        n_new = int(count.item())
        newmom = torch.empty(max(n_new, 1) * 6, dtype=torch.float64, device=mom.device)
        dv.call("cm2_compact_rows_f64", dv.ptr(mom), dv.ptr(old2new), n, 6, dv.ptr(newmom), st)
        self._mom_dev = newmom
        self._old2new_dev = old2new[:n]
        self.__new_npix = n_new
        # O(npix) integer bookkeeping on the host (obspix may be any integer array, or shorter
        # than npix in the reference's default; see process_ces.py:67-68, 346)
        o2n = dv.to_host(self._old2new_dev).astype(np.int64)
        keep = o2n >= 0
        self.old2new = o2n
        self.mask = np.nonzero(keep)[0]
        obspix = np.asarray(self.obspix)
        m = min(len(obspix), n)
        self.obspix = np.concatenate([obspix[:m][keep[:m]], obspix[n:]])

    def _set_obspix(self, new_obspix):
        old2new = np.full(self.oldnpix, -1, dtype=np.int32)
        obspix = np.asarray(self.obspix)
        new_obspix = np.asarray(new_obspix)
        if not (np.all(obspix[:-1] <= obspix[1:]) and np.all(new_obspix[:-1] <= new_obspix[1:])):
            obspix = obspix[np.argsort(obspix, kind="quicksort")]
        idx = np.searchsorted(obspix, new_obspix)
        old2new[idx] = np.arange(len(idx))
        self.old2new = old2new
        self._old2new_dev = dv.to_dev(old2new, torch.int32)
        self.obspix = new_obspix
        self.__new_npix = len(new_obspix)

    def _relabel(self):
        dv.call("cm2_relabel", dv.ptr(self._pix_dev), self.nsamples, dv.ptr(self._old2new_dev), dv.stream())

    def _write_back(self):
        """The reference relabels the caller's ``pixs`` in place (process_ces.py:416)."""
        pixs = self.pixs
        if isinstance(pixs, torch.Tensor):
            pixs.copy_(self._pix_dev.to(pixs.dtype))
            return
        host = dv.to_host(self._pix_dev)
        try:
            pixs[...] = host
        except TypeError:
            for i, v in enumerate(host):
                pixs[i] = int(v)
        dv.register_host_image(pixs, self._pix_dev)

    # ---- reference attribute surface ---------------------------------------------------------
    @property
    def get_new_pixel(self):
        return self.__new_npix, self.obspix

    def _mom_col(self, k):
        key = "mom%d" % k
        if key not in self._host:
            n = self.__new_npix
            self._host[key] = dv.to_host(self._mom_dev.view(-1, 6)[:n, k].contiguous())
        return self._host[key]

    def _need(self, pols, name):
        if self.pol not in pols:
            raise AttributeError("%s is not defined for pol=%d" % (name, self.pol))

    @property
    def counts(self):
        self._need((1, 3), "counts")
        return self._mom_col(0)

    @property
    def cosine(self):
        self._need((3,), "cosine")
        return self._mom_col(1)

    @property
    def sine(self):
        self._need((3,), "sine")
        return self._mom_col(2)

    @property
    def cos2(self):
        self._need((2, 3), "cos2")
        return self._mom_col(3)

    @property
    def sincos(self):
        self._need((2, 3), "sincos")
        return self._mom_col(4)

    @property
    def sin2(self):
        self._need((2, 3), "sin2")
        return self._mom_col(5)

    @property
    def cos(self):
        self._need((2, 3), "cos")
        if "cos" not in self._host:
            self._host["cos"] = dv.to_host(self._cos_dev)
        return self._host["cos"]

    @property
    def sin(self):
        self._need((2, 3), "sin")
        if "sin" not in self._host:
            self._host["sin"] = dv.to_host(self._sin_dev)
        return self._host["sin"]

    def hits(self):
        """Integer hit counts per (new) pixel, bit-exact (int64), from cm2_hits_i64."""
        n = self.__new_npix
        out = torch.empty(max(n, 1), dtype=torch.int64, device=self._pix_dev.device)
        dv.call("cm2_hits_i64", dv.ptr(self._pix_dev), self.nsamples, n, dv.ptr(out), dv.stream())
        return dv.to_host(out[:n])
