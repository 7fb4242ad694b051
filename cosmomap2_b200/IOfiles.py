"""
Input / output of the map-making solve -- the surface of the reference's ``utilities/IOfiles.py``
(same function names, arguments and return tuples), reading and writing the same HDF5 layouts:

* a Constant Elevation Scan pre-processed by the Polarbear AnalysisBackend (IOfiles.py:18-72,
  142-210): ``obspix``, ``n_bolo_pair``, ``n_sample_ces``, ``subscans/{n_sample,t_start}`` and per
  detector pair ``bolo_pair_<i>/{pixel,pol_angle,ground,sum,weight_sum,dif,weight_dif}``;
* the ``system_setup`` test-case layout ``bolo_pair/{pixel,pol_angle,weight,sum}`` (:277-300, 336-349);
* Ritz vectors of the deflation space (:214-256), observed-pixel lists (:258-275), map lists (:351-375).

``h5py`` is used when it is importable; otherwise ``cosmomap2_b200.hdf5lite`` (pure NumPy, classic
HDF5 layout) reads and writes the files.  Host-side only: nothing here touches the GPU; the arrays
returned are what ``ProcessTimeSamples`` / ``SparseLO`` / ``FilterLO`` take.  ``read_ces_shard`` is the
one addition: it reads only the detector pairs of one rank (the TOD is sharded by detector,
distributed.shard_detectors), so an 8-GPU job never holds the whole CES on one host.
"""
from functools import reduce

import numpy as np

from . import hdf5lite

try:                                                       # pragma: no cover - not in this image
    import h5py as _h5
except ImportError:
    _h5 = None


def _open(filename):
    return _h5.File(filename, "r") if _h5 is not None else hdf5lite.File(filename)


def _write(filename, tree):
    """``tree``: nested dict of arrays.  Both back ends store the arrays with their dtype/byte order."""
    if _h5 is None:
        hdf5lite.write(filename, tree)
        return

    def put(group, sub):                                   # pragma: no cover - needs h5py
        for name, val in sub.items():
            if isinstance(val, dict):
                put(group.create_group(name), val)
            else:
                group.create_dataset(name, data=np.asarray(val))
    with _h5.File(filename, "w") as f:                     # pragma: no cover
        put(f, tree)


_I32BE, _F64BE = np.dtype(">i4"), np.dtype(">f8")          # h5t.STD_I32BE / IEEE_F64BE of the reference


def _pair_fields(pol):
    if pol == 1:
        return "sum", "weight_sum"                         # :58-60
    if pol in (2, 3):
        return "dif", "weight_dif"                         # :61-63
    raise RuntimeError("No valid polarization key set!\t=>\tpol=%r" % (pol,))


def _read_pairs(f, pol, pair_ids, subscan=None):
    """Per-pair datasets of one CES file, concatenated detector-major (the TOD layout,
    linearoperators.py:134-140)."""
    dname, wname = _pair_fields(pol)
    pixs, polang, d, weight, ground = [], [], [], [], []
    for i in pair_ids:
        group = f["bolo_pair_" + str(i)]
        pix = np.array(group["pixel"][...])
        if subscan is not None:
            flagging_subscan(pix, subscan)                 # :185-186
        pixs.append(pix)
        polang.append(group["pol_angle"][...])
        ground.append(np.asarray(group["ground"][...]).astype("int"))
        d.append(group[dname][...])
        weight.append(group[wname][...])
    if not pixs:
        z = np.zeros(0)
        return z, np.zeros(0), z, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=int)
    return (np.concatenate(d), np.array(weight), np.concatenate(polang), np.concatenate(pixs),
            np.concatenate(ground))


def read_from_data(filename, pol, npairs=None):
    """One CES file -> ``d, weight, polang, pixs, hp_pixs, ground, n_ces`` (IOfiles.py:18-72)."""
    f = _open(filename)
    try:
        hp_pixs = f["obspix"][...]
        n_bolo_pair = int(f["n_bolo_pair"][...])
        n_ces = f["n_sample_ces"][...]
        n_to_read = n_bolo_pair if npairs is None else npairs
        d, weight, polang, pixs, ground = _read_pairs(f, pol, range(n_to_read))
    finally:
        f.close()
    return d, weight, polang, pixs, hp_pixs, ground, n_ces


def flagging_subscan(unflagged_pix, subscan):
    """Flag (-1), in place, the samples in front of every subscan (IOfiles.py:142-151).  As in the
    reference the samples behind the last subscan stay unflagged."""
    nsamples, tstart = subscan[0], subscan[1]
    k = 0
    for t, n in zip(tstart, nsamples):
        unflagged_pix[k:t] = -1
        k = t + n


def read_from_data_with_subscan_resize(filename, pol, npairs=None):
    """One CES file with its subscan table; samples outside subscans are flagged
    -> ``d, weight, polang, pixs, hp_pixs, ground, n_ces, n_pairs_read, [n_sample, t_start]``
    (IOfiles.py:153-210)."""
    f = _open(filename)
    try:
        hp_pixs = f["obspix"][...]
        n_bolo_pair = int(f["n_bolo_pair"][...])
        n_ces = f["n_sample_ces"][...]
        subscan = [f["subscans/n_sample"][...], f["subscans/t_start"][...]]
        n_to_read = n_bolo_pair if npairs is None else npairs
        d, weight, polang, pixs, ground = _read_pairs(f, pol, range(n_to_read), subscan)
    finally:
        f.close()
    return d, weight, polang, pixs, hp_pixs, ground, n_ces, n_to_read, subscan


def read_ces_shard(filename, pol, rank, world, filtersubscan=True):
    """This rank's detector pairs of one CES file (contiguous, balanced slice of the pairs):
    the same tuple as ``read_from_data_with_subscan_resize`` (or ``read_from_data`` without the last
    two entries' subscan table when ``filtersubscan`` is False), with ``n_pairs_read`` = the number of
    local pairs.  Concatenating the shards of all ranks in rank order gives the unsharded arrays."""
    from .distributed import shard_detectors
    f = _open(filename)
    try:
        hp_pixs = f["obspix"][...]
        n_bolo_pair = int(f["n_bolo_pair"][...])
        n_ces = f["n_sample_ces"][...]
        lo, hi = shard_detectors(n_bolo_pair, world, rank)
        subscan = [f["subscans/n_sample"][...], f["subscans/t_start"][...]] if filtersubscan else None
        d, weight, polang, pixs, ground = _read_pairs(f, pol, range(lo, hi), subscan)
    finally:
        f.close()
    if filtersubscan:
        return d, weight, polang, pixs, hp_pixs, ground, n_ces, hi - lo, subscan
    return d, weight, polang, pixs, hp_pixs, ground, n_ces


def read_multiple_ces(filelist, pol, npairs=None, filtersubscan=True):
    """Several CES files concatenated CES-major (IOfiles.py:73-121).  With ``filtersubscan`` the
    per-CES subscan tables, samples per pair and pairs per CES come back as lists -- exactly the
    arguments ``FilterLO`` takes.  ``hp_pixs`` is the one of the LAST file, as in the reference (:112)."""
    readf = read_from_data_with_subscan_resize if filtersubscan else read_from_data
    subscan, tstart, bolopairs_per_ces, samples_per_bolopair = [], [], [], []
    pixs, polang, d, weight, ground = [], [], [], [], []
    outdata = None
    for fname in filelist:
        outdata = readf(fname, pol, npairs=npairs)
        d.append(outdata[0])
        weight.append(outdata[1])
        polang.append(outdata[2])
        pixs.append(outdata[3])
        ground.append(outdata[5])
        if filtersubscan:
            samples_per_bolopair.append(outdata[6])
            bolopairs_per_ces.append(outdata[7])
            subscan.append(outdata[8][0])
            tstart.append(outdata[8][1])
    hp_pixs = [outdata[4]]
    head = (np.concatenate(d), np.concatenate(weight), np.concatenate(polang), np.concatenate(pixs),
            np.concatenate(hp_pixs), np.concatenate(ground))
    if filtersubscan:
        return head + (subscan, tstart, samples_per_bolopair, bolopairs_per_ces)
    return head + (samples_per_bolopair, bolopairs_per_ces)


def flagging_not_in_allCES(CES_pixs):
    """Flag (-1), in place, the pixels that are not observed by every CES (IOfiles.py:123-138)."""
    inters = reduce(np.intersect1d, CES_pixs)
    for pixs in CES_pixs:
        pixs[np.isin(pixs, inters, invert=True)] = -1


# ---- artefacts of the solve ---------------------------------------------------------------------
def write_ritz_eigenvectors_to_hdf5(z, filename, eigvals=None):
    """Deflation space ``Z`` (n x r) [+ Ritz values] (IOfiles.py:214-238): a checkpoint that lets a
    re-run skip the Arnoldi phase.  ``z`` may be a CUDA tensor."""
    z = _to_host(z)
    if np.iscomplexobj(z):
        raise NotImplementedError("complex Ritz vectors (variable-length HDF5 type) are not supported")
    tree = {"Ritz_eigenvectors": {"n_eigenvectors": np.asarray(z.shape[1], dtype=_I32BE),
                                  "Eigenvectors": np.asarray(z, dtype=np.float64)}}
    if eigvals is not None:
        tree["Ritz_eigenvalues"] = np.asarray(_to_host(eigvals), dtype=np.float64)
    _write(filename, tree)


def read_ritz_eigenvectors_from_hdf5(filename, eigvals=False):
    """-> ``z, n_eigenvals[, eigenvals]`` (IOfiles.py:240-256)."""
    f = _open(filename)
    try:
        n_eigenvals = f["Ritz_eigenvectors/n_eigenvectors"][...]
        z = f["Ritz_eigenvectors/Eigenvectors"][...]
        if eigvals:
            return z, n_eigenvals, f["Ritz_eigenvalues"][...]
        return z, n_eigenvals
    finally:
        f.close()


def read_obspix_from_hdf5(filename):
    f = _open(filename)
    try:
        return f["obspix"][...]                            # :258-267
    finally:
        f.close()


def write_obspix_to_hdf5(filename, obspix):
    _write(filename, {"obspix": np.asarray(obspix, dtype=_I32BE)})   # :268-275


def write_to_hdf5(filename, obs_pixels, noise_values, d, phi=None):
    """The ``system_setup`` test-case layout (IOfiles.py:277-300): big-endian int32 pixels, big-endian
    fp64 weight / sum / pol_angle under the group ``bolo_pair``."""
    group = {"pixel": np.asarray(obs_pixels, dtype=_I32BE), "weight": np.asarray(noise_values, dtype=_F64BE),
             "sum": np.asarray(d, dtype=_F64BE)}
    if phi is not None:
        group["pol_angle"] = np.asarray(phi, dtype=_F64BE)
    _write(filename, {"bolo_pair": group})


def read_from_hdf5(filename):
    """-> ``det, obs_pix, polang, weight`` (IOfiles.py:336-349)."""
    f = _open(filename)
    try:
        return (f["bolo_pair/sum"][...], f["/bolo_pair/pixel"][...], f["/bolo_pair/pol_angle"][...],
                f["/bolo_pair/weight"][...])
    finally:
        f.close()


def save_maplist(maplist, filename):
    """The solution at every iteration step (IOfiles.py:351-362)."""
    tree = {"Nmaps": np.asarray(len(maplist), dtype=_I32BE)}
    for i, m in enumerate(maplist):
        tree["Map" + str(i)] = np.asarray(_to_host(m), dtype=_F64BE)
    _write(filename, tree)


def read_maplist(filename):
    """-> ``maps, nmaps`` (IOfiles.py:364-375)."""
    f = _open(filename)
    try:
        nmaps = f["Nmaps"][...]
        return [np.array(f["Map" + str(i)][...]).T for i in range(int(nmaps))], nmaps
    finally:
        f.close()


def full2cutskymap(hp_map, pol, npix, observpix):
    """Full-sky HEALPix maps ``[I, Q, U]`` -> the interleaved cut-sky vector the operators use
    (IOfiles.py:377-393); the inverse of ``reorganize_map``.  ``pol = 1`` returns the list of observed
    maps, as the reference does."""
    obsmap = [np.asarray(m)[observpix] for m in hp_map]
    if pol == 1:
        return obsmap
    x = np.zeros(pol * npix)
    for k in range(pol):
        x[k::pol] = obsmap[k][:npix]
    return x


def obspix2mask(obspix, nside, fname=None):
    """utilities/healpy_functions.py:22-46 without healpy: 1 on the observed pixels of a
    12 nside^2 map."""
    if fname is not None:
        raise NotImplementedError("writing HEALPix FITS files needs healpy; pass fname=None")
    mask = np.zeros(12 * int(nside) * int(nside))
    mask[np.asarray(obspix)] = 1
    return mask


def find_common_obspix(nside, pathtofiles, n_files):
    """Intersection of the observed pixels of ``obspix_<k>.hdf5``, k < n_files, written to
    ``common_obspix.hdf5``; also the coverage map sum_k (1+k) mask_k (IOfiles.py:395-409).  The common
    pixel set is what ``ProcessTimeSamples(..., obspix2=...)`` takes in multi-CES / multi-GPU runs."""
    mask = 0.
    obspix_set = []
    for offset in range(n_files):
        hp_pixs = read_obspix_from_hdf5(pathtofiles + "obspix_" + str(offset) + ".hdf5")
        obspix_set.append(hp_pixs)
        mask = mask + (1 + offset) * obspix2mask(hp_pixs, nside)
    common_obsp = reduce(np.intersect1d, obspix_set)
    write_obspix_to_hdf5(pathtofiles + "common_obspix.hdf5", common_obsp)
    return obspix_set, mask


# ---- synthetic CES files (no counterpart in the reference, which only reads this schema) ---------------
def write_ces_to_hdf5(filename, obspix, pixel, pol_angle, ground, n_sample_ces, subscan_nsample, subscan_tstart,
                      sum_=None, weight_sum=None, dif=None, weight_dif=None):
    """Write one CES in the AnalysisBackend schema ``read_from_data*`` expect.  ``pixel``, ``pol_angle``,
    ``ground`` (and the data streams) are lists with one array per detector pair; ``weight_*`` one scalar
    per pair."""
    npair = len(pixel)
    tree = {"obspix": np.asarray(obspix, dtype=_I32BE), "n_bolo_pair": np.asarray(npair, dtype=_I32BE),
            "n_sample_ces": np.asarray(n_sample_ces, dtype=_I32BE),
            "subscans": {"n_sample": np.asarray(subscan_nsample, dtype=np.int64),
                         "t_start": np.asarray(subscan_tstart, dtype=np.int64)}}
    for i in range(npair):
        g = {"pixel": np.asarray(pixel[i], dtype=_I32BE), "pol_angle": np.asarray(pol_angle[i], dtype=_F64BE),
             "ground": np.asarray(ground[i], dtype=_I32BE)}
        for name, val in (("sum", sum_), ("weight_sum", weight_sum), ("dif", dif), ("weight_dif", weight_dif)):
            if val is not None:
                g[name] = np.asarray(val[i], dtype=_F64BE)
        tree["bolo_pair_" + str(i)] = g
    _write(filename, tree)


def _to_host(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
    except ImportError:                                    # pragma: no cover
        pass
    return np.asarray(a)
