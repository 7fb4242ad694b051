"""CPU tests of the input / output row: the HDF5 reader against the
reference's OWN fixtures (data/testcase_block_diag_{3,4}.hdf5, written by h5py through the reference's
write_to_hdf5 -- copies under tests/golden/), writer -> reader round trips in every layout the reader
supports, and the reference's reader functions on synthetic CES files in the AnalysisBackend schema."""
import os

import numpy as np
import pytest

from cosmomap2_b200 import IOfiles as io
from cosmomap2_b200 import hdf5lite

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,pixscale,wshape", [("testcase_block_diag_3.hdf5", 3, (2, 2)),
                                                  ("testcase_block_diag_4.hdf5", 1, (2,))])
def test_reader_decodes_the_reference_fixtures(name, pixscale, wshape):
    """Files written by the reference (h5py, big-endian types): system_setup(nt=100, npix=15, nb=2)."""
    det, pix, phi, weight = io.read_from_hdf5(os.path.join(GOLDEN, name))
    assert det.shape == (100,) and pix.shape == (100,) and phi.shape == (100,) and weight.shape == wshape
    assert pix.dtype == np.int32 and det.dtype == np.float64          # native byte order, like h5py
    assert pix.min() == 0 and pix.max() == 14 * pixscale and np.all(pix % pixscale == 0)
    # angles_gen: theta0 + 2 pi 2.5/200 i  (utilities_functions.py:99-107)
    assert np.allclose(np.diff(phi), 2 * np.pi * 2.5 / 200., rtol=0, atol=1e-14)
    assert 0.0 < det.min() and det.max() < 1.0 and np.all(weight > 0) and np.all(weight < 1)
    f = hdf5lite.File(os.path.join(GOLDEN, name))
    assert f.keys() == ["bolo_pair"] and sorted(f["bolo_pair"].keys()) == ["pixel", "pol_angle", "sum", "weight"]
    assert f["bolo_pair"]["pixel"].dtype == np.dtype(">i4") and f["/bolo_pair/sum"].dtype == np.dtype(">f8")
    with pytest.raises(KeyError):
        f["bolo_pair/nothing"]


def test_write_to_hdf5_round_trip_equals_the_fixture(tmp_path):
    """write_to_hdf5 (IOfiles.py:277-300) -> read_from_hdf5 gives back the reference's fixture content,
    with the same on-disk types."""
    det, pix, phi, weight = io.read_from_hdf5(os.path.join(GOLDEN, "testcase_block_diag_3.hdf5"))
    out = str(tmp_path / "again.hdf5")
    io.write_to_hdf5(out, pix, weight, det, phi=phi)
    det2, pix2, phi2, weight2 = io.read_from_hdf5(out)
    for a, b in ((det, det2), (pix, pix2), (phi, phi2), (weight, weight2)):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    f = hdf5lite.File(out)
    assert f["bolo_pair/pixel"].dtype == np.dtype(">i4") and f["bolo_pair/weight"].dtype == np.dtype(">f8")


def test_layouts_types_and_many_members(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 5))
    tree = {"scalar": np.asarray(7, dtype=">i4"), "u8": np.arange(5, dtype=np.uint8), "f4": np.float32([1.5, -2.25]),
            "le": np.arange(12, dtype="<i8").reshape(3, 4), "empty": np.zeros(0),
            "grp": {"chunked": hdf5lite.Chunked(a, (10, 2)),
                    "packed": hdf5lite.Chunked(a, (16, 5), deflate=4, shuffle=True),
                    "ints": hdf5lite.Chunked(np.arange(1000, dtype=">i4"), (300,), deflate=1),
                    "deep": {"x": np.float64([3.0])}}}
    for i in range(30):                                    # more than one symbol-table node
        tree["grp"]["m%02d" % i] = np.full(3, float(i))
    p = str(tmp_path / "t.h5")
    hdf5lite.write(p, tree)
    with hdf5lite.File(p) as f:
        assert int(f["scalar"][...]) == 7 and f["scalar"][...].shape == ()
        assert np.array_equal(f["u8"][...], tree["u8"]) and np.array_equal(f["f4"][...], tree["f4"])
        assert np.array_equal(f["le"][...], tree["le"]) and f["empty"][...].shape == (0,)
        for k in ("chunked", "packed"):
            assert np.array_equal(f["grp"][k][...], a)
        assert np.array_equal(f["grp/ints"][...], np.arange(1000))
        assert f["grp/deep/x"][...][0] == 3.0 and len(f["grp"].keys()) == 34
        assert np.array_equal(f["grp/m17"][...], np.full(3, 17.0))
        assert np.array_equal(f["grp/chunked"][3:5, 1], a[3:5, 1])
    with open(p, "r+b") as fh:                             # a libver='latest' superblock is refused, loudly
        fh.seek(8)
        fh.write(b"\x02")
    with pytest.raises(NotImplementedError):
        hdf5lite.File(p)
    with pytest.raises(hdf5lite.Hdf5Error):
        open(p, "wb").write(b"not hdf5 at all")
        hdf5lite.File(p)


def _make_ces(tmp_path, name, rng, npair, ns, nsub, pol_fields=True):
    cuts = np.sort(rng.choice(np.arange(1, ns), size=2 * nsub, replace=False))
    t_start = cuts[0::2].astype(np.int64)
    n_sample = (cuts[1::2] - cuts[0::2]).astype(np.int64)
    obspix = np.sort(rng.choice(12 * 16 * 16, 40, replace=False))
    pixel = [rng.integers(0, 40, size=ns) for _ in range(npair)]
    phi = [rng.uniform(0, np.pi, size=ns) for _ in range(npair)]
    ground = [rng.integers(-1, 20, size=ns) for _ in range(npair)]
    s = [rng.standard_normal(ns) for _ in range(npair)]
    dd = [rng.standard_normal(ns) for _ in range(npair)]
    ws, wd = rng.uniform(0.5, 2, npair), rng.uniform(0.5, 2, npair)
    path = str(tmp_path / name)
    io.write_ces_to_hdf5(path, obspix, pixel, phi, ground, ns, n_sample, t_start, sum_=s, weight_sum=ws, dif=dd,
                         weight_dif=wd)
    return path, dict(obspix=obspix, pixel=pixel, phi=phi, ground=ground, s=s, d=dd, ws=ws, wd=wd, ns=ns,
                      n_sample=n_sample, t_start=t_start, npair=npair)


def _flagged(pix, n_sample, t_start):
    """Direct restatement of flagging_subscan (IOfiles.py:142-151) with a mask."""
    keep = np.zeros(len(pix), dtype=bool)
    for t, n in zip(t_start, n_sample):
        keep[t:t + n] = True
    keep[t_start[-1] + n_sample[-1]:] = True               # the tail behind the last subscan is NOT flagged
    out = pix.copy()
    out[~keep] = -1
    return out


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_read_from_data_and_subscan_flagging(tmp_path, pol):
    rng = np.random.default_rng(pol)
    path, c = _make_ces(tmp_path, "ces.hdf5", rng, npair=3, ns=500, nsub=4)
    d, weight, polang, pixs, hp_pixs, ground, n_ces = io.read_from_data(path, pol)
    data, w = (c["s"], c["ws"]) if pol == 1 else (c["d"], c["wd"])
    assert np.array_equal(d, np.concatenate(data)) and np.array_equal(weight, w)
    assert np.array_equal(polang, np.concatenate(c["phi"])) and np.array_equal(pixs, np.concatenate(c["pixel"]))
    assert np.array_equal(hp_pixs, c["obspix"]) and np.array_equal(ground, np.concatenate(c["ground"]))
    assert int(n_ces) == 500 and ground.dtype.kind == "i"
    # npairs limits the pairs read; the subscan variant flags the samples outside subscans
    out = io.read_from_data_with_subscan_resize(path, pol, npairs=2)
    assert out[7] == 2 and len(out[3]) == 1000
    assert np.array_equal(out[8][0], c["n_sample"]) and np.array_equal(out[8][1], c["t_start"])
    want = np.concatenate([_flagged(c["pixel"][i], c["n_sample"], c["t_start"]) for i in range(2)])
    assert np.array_equal(out[3], want)
    with pytest.raises(RuntimeError):
        io.read_from_data(path, 4)


def test_read_multiple_ces_feeds_filterlo_and_shards_concatenate(tmp_path):
    import oracle
    rng = np.random.default_rng(7)
    p0, c0 = _make_ces(tmp_path, "ces0.hdf5", rng, npair=3, ns=400, nsub=3)
    p1, c1 = _make_ces(tmp_path, "ces1.hdf5", rng, npair=2, ns=300, nsub=4)
    d, weight, polang, pixs, hp_pixs, ground, subscan, tstart, ns_list, nb_list = io.read_multiple_ces([p0, p1], 3)
    assert len(d) == 3 * 400 + 2 * 300 and [int(n) for n in ns_list] == [400, 300] and nb_list == [3, 2]
    assert np.array_equal(weight, np.concatenate([c0["wd"], c1["wd"]]))
    assert np.array_equal(hp_pixs, c1["obspix"])           # the LAST file's, as in the reference (:112)
    want = np.concatenate([_flagged(c["pixel"][i], c["n_sample"], c["t_start"]) for c in (c0, c1)
                           for i in range(c["npair"])])
    assert np.array_equal(pixs, want)
    # the lists are exactly FilterLO's arguments (linearoperators.py:263-275)
    F = oracle.FilterLO(len(d), [subscan, tstart], ns_list, nb_list, pixs)
    y = F * d
    t0, n0 = int(c0["t_start"][1]), int(c0["n_sample"][1])
    seg = slice(400 + t0, 400 + t0 + n0)                   # CES 0, pair 1, subscan 1
    good = pixs[seg] != -1
    assert abs(np.mean(y[seg][good])) < 1e-12
    without = io.read_multiple_ces([p0, p1], 3, filtersubscan=False)
    assert len(without) == 8 and np.array_equal(without[3], np.concatenate(c0["pixel"] + c1["pixel"]))
    # detector sharding: the shards of all ranks, in rank order, are the unsharded arrays
    full = io.read_from_data_with_subscan_resize(p0, 3)
    for world in (2, 3, 4):
        parts = [io.read_ces_shard(p0, 3, r, world) for r in range(world)]
        assert sum(p[7] for p in parts) == 3
        for k in (0, 2, 3, 5):
            assert np.array_equal(np.concatenate([p[k] for p in parts]), full[k])
        assert np.array_equal(np.concatenate([np.atleast_1d(p[1]) for p in parts if p[7]]), full[1])


def test_artefacts_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    z = rng.standard_normal((60, 4))
    vals = rng.uniform(0, 1e-2, 4)
    p = str(tmp_path / "ritz.hdf5")
    io.write_ritz_eigenvectors_to_hdf5(z, p, eigvals=vals)
    z2, n2, v2 = io.read_ritz_eigenvectors_from_hdf5(p, eigvals=True)
    assert np.array_equal(z, z2) and int(n2) == 4 and np.array_equal(vals, v2)
    assert len(io.read_ritz_eigenvectors_from_hdf5(p)) == 2
    maps = [rng.standard_normal(9), rng.standard_normal((3, 2))]
    io.save_maplist(maps, str(tmp_path / "maps.hdf5"))
    m2, nm = io.read_maplist(str(tmp_path / "maps.hdf5"))
    assert int(nm) == 2 and np.array_equal(m2[0], maps[0]) and np.array_equal(m2[1], maps[1].T)
    # obspix files and their intersection (find_common_obspix, IOfiles.py:395-409)
    sets = [np.array([1, 5, 9, 40, 41]), np.array([5, 9, 40, 700]), np.array([0, 5, 40, 41])]
    for k, s in enumerate(sets):
        io.write_obspix_to_hdf5(str(tmp_path / ("obspix_%d.hdf5" % k)), s)
    got, mask = io.find_common_obspix(8, str(tmp_path) + os.sep, 3)
    assert all(np.array_equal(a, b) for a, b in zip(got, sets))
    assert np.array_equal(io.read_obspix_from_hdf5(str(tmp_path / "common_obspix.hdf5")), [5, 40])
    assert mask[5] == 6 and mask[41] == 4 and mask[700] == 2 and mask[2] == 0 and len(mask) == 768
    # flagging_not_in_allCES (in place)
    ces = [np.array([1, 2, 3, 4]), np.array([2, 4, 6])]
    io.flagging_not_in_allCES(ces)
    assert ces[0].tolist() == [-1, 2, -1, 4] and ces[1].tolist() == [2, 4, -1]


def test_full2cutskymap_inverts_reorganize_map():
    import oracle
    rng = np.random.default_rng(5)
    nside, npix = 4, 20
    obspix = np.sort(rng.choice(12 * nside * nside, npix, replace=False))
    for pol in (2, 3):
        m = rng.standard_normal(pol * npix)
        full = oracle.reorganize_map(m, obspix, npix, nside, pol)
        assert np.array_equal(io.full2cutskymap(full, pol, npix, obspix), m)
    one = oracle.reorganize_map(rng.standard_normal(npix), obspix, npix, nside, 1)
    assert len(io.full2cutskymap(one, 1, npix, obspix)) == 1
