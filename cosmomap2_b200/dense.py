"""
Host side of the dense tall-skinny contractions of the deflation path (csrc/dense_z.cu, fp64 tensor
cores): ``gram(Xt, Yt) = X^T Y`` and ``combine(Vt, U) = (V U)^T``.  Tall matrices are handled in the
layout the kernels read: a ``(r, n)`` row-contiguous CUDA tensor IS the column-major ``n x r`` matrix
with leading dimension n ("columns-as-rows": ``Zt[j]`` is column j of Z).
"""
import numpy as np
import torch

from . import _device as dv

_work = {}


def _gram_work():
    dev = torch.cuda.current_device()
    if dev not in _work:
        _work[dev] = dv.empty_f64(int(dv.call("cm2_dense_gram_work_doubles")))
    return _work[dev]


def _rows(t):
    """(r, n) CUDA fp64 tensor whose rows are contiguous with a common stride -> (tensor, ld)."""
    t = dv.to_dev_f64(t)
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return t, int(ld)


def gram(Xt, Yt):
    """``X^T Y`` (r1 x r2 CUDA tensor) for ``Xt`` (r1, n), ``Yt`` (r2, n) -- i.e. ``Xt @ Yt.T`` -- in one
    pass over both (cm2_dense_gram); the reference's ``dgemm(Z, Az.T)`` (linearoperators.py:1019)."""
    Xt, ldx = _rows(Xt)
    Yt, ldy = _rows(Yt)
    r1, n = Xt.shape
    r2 = Yt.shape[0]
    if Yt.shape[1] != n:
        raise ValueError("gram: row counts differ")
    out = dv.empty_f64(r1 * r2)
    dv.call("cm2_dense_gram", dv.ptr(Xt), ldx, int(r1), dv.ptr(Yt), ldy, int(r2), int(n), dv.ptr(out), int(r1),
            dv.ptr(_gram_work()), dv.stream())
    return out.view(r2, r1).t()                 # column-major r1 x r2


def combine(Vt, U):
    """``(V U)^T`` as an (r, n) CUDA tensor for ``Vt`` (m, n) and ``U`` (m, r) host or device
    (cm2_dense_combine): the Ritz-vector assembly ``Z = V[:, :m] U`` (deflationlib.py:204-219)."""
    Vt, ldv = _rows(Vt)
    m, n = Vt.shape
    Uh = U if isinstance(U, torch.Tensor) else np.asarray(U, dtype=np.float64)
    if Uh.shape[0] != m:
        raise ValueError("combine: U must have one row per basis vector")
    r = int(Uh.shape[1])
    # column-major m x r
    Ud = dv.to_dev_f64(Uh.t().contiguous() if isinstance(Uh, torch.Tensor) else np.ascontiguousarray(Uh.T))
    npad = (n + 3) // 4 * 4                     # leading dimension a multiple of 4: 256-bit stores
    Zbuf = dv.empty_f64(max(r * npad, 1))
    dv.call("cm2_dense_combine", dv.ptr(Vt), ldv, int(n), int(m), dv.ptr(Ud), int(m), r, dv.ptr(Zbuf), int(npad),
            dv.stream())
    return Zbuf.view(r, npad)[:, :n]
