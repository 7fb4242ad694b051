"""The workloads behind bench.py's `secondary` block and the examples (cosmomap2_b200/workloads.py), at sizes that run
in seconds: the functions return what the bench line promises and the numbers are sane."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wl():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cosmomap2_b200 import workloads
    return workloads


def _roofline_ok(r):
    for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "algorithmic_bytes_per_launch", "kernel_ms"):
        assert k in r
    assert r["achieved"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12


def test_correlated_workload(wl):
    out = wl.correlated(nt=4e5, ndet=4, nband=300, nside=128, nx=100, ny=60, rtol=1e-6, maxiter=400, time_iters=2,
                        two_level_r=8)
    assert out["cg"]["info"] == 0 and out["cg"]["true_relres"] < 5e-6
    m2 = out["M_2lvl_scan_space"]
    assert m2["info"] == 0 and m2["iterations"] <= out["cg"]["iterations"] and m2["Ax_agreement"] < 1e-4
    assert out["plan"][-1] == "_FusedFilterP" and out["nband"] == 300
    assert out["symmetry"]["rel_to_norms"] < 1e-12
    _roofline_ok(out["roofline"])
    _roofline_ok(out["roofline_A_apply"])


def test_two_level_workload_scan_space_pays(wl):
    out = wl.two_level(nt=1.6e6, nside=128, nx=120, ny=96, ndet=8, r=12, coarse="scan", smooth=2, rtol=1e-8, maxiter=2000,
                       time_iters=2)
    bd, m2 = out["M_BD"], out["M_2lvl"]
    assert bd["info"] == 0 and m2["info"] == 0
    assert m2["iterations"] <= 0.6 * bd["iterations"], (bd["iterations"], m2["iterations"])
    assert out["Ax_agreement"] < 1e-6 and out["deflation"]["kind"] == "scan"
    assert out["M_2lvl_apply_ms"] > 0
    _roofline_ok(out["roofline"])


def test_two_level_workload_ritz_route(wl):
    out = wl.two_level(nt=4e5, nside=128, nx=60, ny=24, ndet=4, r=8, coarse="ritz", arnoldi=60, rtol=1e-8, maxiter=2000,
                       time_iters=2)
    assert out["M_BD"]["info"] == 0 and out["M_2lvl"]["info"] == 0 and out["deflation"]["kind"] == "ritz"
    assert out["deflation"]["arnoldi_steps"] > 8


def test_white_workload(wl):
    out = wl.white(nt=1e6, nside=128, nx=120, ny=80, ndet=8, steps=3)
    assert out["cg_info"] == 0 and out["relres"] < 1e-12 and out["solver"] == "PCG"
    _roofline_ok(out["roofline"])


def test_pattern_generators(wl):
    import torch
    nt, ns, pix, phi, sl, ss, g = wl.make_scan(200000, 64, 80, 40, 4, 8.0, seed=1, tilt_deg=30.0)
    p = pix.cpu().numpy()
    good = p >= 0
    assert nt == 200000 and ns == 50000 and good.mean() > 0.9
    rows = p[good] // 256
    assert len(np.unique(rows)) == 40                           # the whole patch height is covered
    # a tilted sweep changes row every few pixels: most pixel changes are not to the neighbouring index
    d = np.diff(p[good].astype(np.int64))
    ch = d[d != 0]
    assert 0.2 < np.mean(np.abs(ch) == 1) < 0.8
    pr, phr, _ = wl.random_pointing(100000, 64, 80, 40, seed=2)
    assert pr.dtype == torch.int32 and int(pr.min()) >= 0 and len(torch.unique(pr)) > 3000
