"""
Krylov / deflation helpers with the names and signatures of interfaces/deflationlib.py, running on
device vectors.

* ``arnoldi``, ``build_hess``, ``build_Z``: the in-tree Euclidean Arnoldi (deflationlib.py:17-184),
  restated literally (modified Gram-Schmidt, the element-wise stop test of :101, ``RuntimeError`` at
  ``inner_m``); ``build_Z`` uses eigenvector COLUMNS (the reference indexes rows and transposes a
  list, :172-174, 183 -- unreachable as written; documented deviation).
* ``run_krypy_arnoldi``, ``find_ritz_eigenvalues`` (:187-219): the reference delegates to the
  third-party ``krypy`` (absent, unpinned).  Restated from its published algorithm: Arnoldi in the
  M^-1 inner product (V = M P, V^T P = I), Ritz pairs of the Hermitian part of H.  Deviation,
  recorded in DESIGN.md: the default orthogonalisation is two-pass block Gram-Schmidt (krypy's
  'dmgs'); the one-pass 'mgs' krypy defaults to loses bi-orthogonality and manufactures spurious
  theta ~ 0 Ritz values that the theta < 1e-2 rule would select.  ``ortho='mgs'`` is available.
  The block passes are two tall-skinny device kernels (V^T w, w -= P h) per pass.
"""
import numpy as np
import torch
from scipy.linalg import eigh

from . import _device as dv
from . import linop as lp
from .utilities import dgemm, norm2  # noqa: F401


def _dot(a, b, out):
    dv.call("cm2_dot", dv.ptr(a), dv.ptr(b), a.numel(), dv.ptr(out), dv.stream())
    return float(out.item())


def _as_op(A):
    if isinstance(A, lp.LinearOperator):
        return A
    return lp.LinearOperator(A.shape[1], A.shape[0], matvec=A.matvec, device=False)


def arnoldi(A, b, x0=None, tol=1e-5, maxiter=1000, inner_m=30):
    """interfaces/deflationlib.py:17-113.  Returns ``(vs, hs, m)``: ``vs`` the orthonormal basis
    (list of vectors), ``hs`` the Hessenberg columns (column j has j+2 entries)."""
    host = not isinstance(b, torch.Tensor)
    if host and not np.isfinite(b).all():
        raise ValueError("RHS must contain only finite numbers")
    A = _as_op(A)
    bd = dv.to_dev_f64(b)
    if not host and not bool(torch.isfinite(bd).all().item()):
        raise ValueError("RHS must contain only finite numbers")
    n = bd.numel()
    s = dv.zeros_f64(1)
    b_norm = np.sqrt(_dot(bd, bd, s))
    if b_norm == 0:
        b_norm = 1
    r_outer = bd.clone()
    ax = A._apply(dv.to_dev_f64(x0))
    dv.call("cm2_axpby", -1.0, dv.ptr(ax), 1.0, dv.ptr(r_outer), n, dv.stream())
    r_norm = np.sqrt(_dot(r_outer, r_outer, s))
    if r_norm < tol * b_norm or r_norm < tol:
        return None, None, 0
    dv.call("cm2_axpby", 0.0, dv.ptr(r_outer), 1.0 / r_norm, dv.ptr(r_outer), n, dv.stream())
    vs = [r_outer]
    hs = []

    def ret(vs):
        return [dv.to_host(v) for v in vs] if host else vs

    for j in range(1, 1 + inner_m):
        v_new = A._apply(vs[j - 1])
        if v_new.data_ptr() == vs[j - 1].data_ptr():
            v_new = v_new.clone()
        hcur = []
        for v in vs:                                     # modified Gram-Schmidt (:92-96)
            alpha = _dot(v, v_new, s)
            hcur.append(alpha)
            dv.call("cm2_axpby", -alpha, dv.ptr(v), 1.0, dv.ptr(v_new), n, dv.stream())
        hcur.append(np.sqrt(_dot(v_new, v_new, s)))
        dv.call("cm2_axpby", 0.0, dv.ptr(v_new), 1.0 / hcur[-1], dv.ptr(v_new), n, dv.stream())
        if j >= n:
            raise IndexError("index %d is out of bounds for axis 0 with size %d" % (j, n))
        if abs(float(v_new[j].item()) * hcur[-1]) <= tol:   # the reference's stop test (:101)
            hs.append(hcur)
            return ret(vs), hs, j
        vs.append(v_new)
        hs.append(hcur)
        if j == inner_m:
            raise RuntimeError("Convergence not achieved within the Arnoldi algorithm")


def build_hess(h, m):
    """interfaces/deflationlib.py:115-137."""
    hess = np.zeros((m, m))
    for q in range(m - 1):
        hess[:(q + 2), q] = h[q]
    hess[:m, m - 1] = h[-1][:m]
    return hess


def build_Z(z, y, w, eps):
    """interfaces/deflationlib.py:140-184 (eigenvector columns; see module docstring)."""
    m = len(z)
    sel = [i for i in range(m) if abs(z[i]) <= eps]
    r = len(sel)
    if r == 0:
        raise RuntimeError("No Ritz eigenvalue are found smaller than fixed threshold %.1g " % eps)
    U = np.asarray(y)[:, sel]
    if isinstance(w[0], torch.Tensor):
        W = torch.stack(list(w[:m]))                           # m x n
        return torch.matmul(dv.to_dev_f64(np.ascontiguousarray(U.T)), W).t(), r
    W = np.asarray(w)[:m]
    return dv.to_host(torch.matmul(dv.to_dev_f64(np.ascontiguousarray(U.T)), dv.to_dev_f64(W))).T, r


def krypy_arnoldi(A, x0, M=None, maxiter=None, ortho="dmgs", tol_invariant=1e-14):
    """Preconditioned Arnoldi with krypy's semantics: returns ``V (n x (m+1)), H ((m+1) x m),
    P (n x (m+1))`` with ``M A V_m = V_{m+1} H`` and ``V^T P = I``.  Tensors if ``x0`` is a CUDA
    tensor, NumPy arrays otherwise."""
    host = not isinstance(x0, torch.Tensor)
    A = _as_op(A)
    M = _as_op(M) if M is not None else None
    p0 = dv.to_dev_f64(x0).reshape(-1)
    n = p0.numel()
    maxiter = n if maxiter is None else min(int(maxiter), n)
    st = dv.stream
    s = dv.zeros_f64(1)
    V = torch.zeros((maxiter + 1, n), dtype=torch.float64, device=p0.device)
    P = torch.zeros((maxiter + 1, n), dtype=torch.float64, device=p0.device) if M is not None else V
    H = np.zeros((maxiter + 1, maxiter))
    v0 = M._apply(p0) if M is not None else p0
    nrm = np.sqrt(_dot(p0, v0, s))
    V[0].copy_(v0)
    V[0].mul_(1.0 / nrm)
    if M is not None:
        P[0].copy_(p0)
        P[0].mul_(1.0 / nrm)
    work = dv.empty_f64(int(dv.call("cm2_defl_work_doubles", int(maxiter + 1))))
    hdev = dv.empty_f64(maxiter + 1)
    k = 0
    invariant = False
    while k < maxiter and not invariant:
        w = A._apply(V[k])
        if w.data_ptr() == V[k].data_ptr():
            w = w.clone()
        if ortho == "mgs":
            for j in range(k + 1):
                a = _dot(V[j], w, s)
                H[j, k] += a
                dv.call("cm2_axpby", -a, dv.ptr(P[j]), 1.0, dv.ptr(w), n, st())
        else:
            for _ in range(2):                                       # block Gram-Schmidt, two passes
                dv.call("cm2_defl_zt_apply", dv.ptr(V), n, k + 1, n, dv.ptr(w), 1, n, dv.ptr(hdev), dv.ptr(work), st())
                dv.call("cm2_defl_z_apply", dv.ptr(P), n, k + 1, n, dv.ptr(hdev), -1.0, 1.0, dv.ptr(w), dv.ptr(w), st())
                H[:k + 1, k] += dv.to_host(hdev[:k + 1])
        Mw = M._apply(w) if M is not None else w
        hk = np.sqrt(abs(_dot(w, Mw, s)))
        H[k + 1, k] = hk
        if hk / np.abs(H[:k + 2, :k + 1]).max() <= tol_invariant:
            invariant = True
        else:
            if M is not None:
                P[k + 1].copy_(w)
                P[k + 1].mul_(1.0 / hk)
            V[k + 1].copy_(Mw)
            V[k + 1].mul_(1.0 / hk)
        k += 1
    if invariant:
        Vo, Ho, Po = V[:k], H[:k, :k], P[:k]
    else:
        Vo, Ho, Po = V[:k + 1], H[:k + 1, :k], P[:k + 1]
    if host:
        return dv.to_host(Vo).T, Ho, dv.to_host(Po).T
    return Vo.t(), Ho, Po.t()


def run_krypy_arnoldi(A, x0, M, tol, maxiter=None, ortho="dmgs"):
    """interfaces/deflationlib.py:187-202 -> ``(v, h, m)``."""
    v, h, p = krypy_arnoldi(A, x0, M=M, maxiter=maxiter, ortho=ortho)
    return v, h, v.shape[1]


def krypy_ritz(H, V=None, hermitian=True):
    """krypy.utils.ritz(H, V=V, hermitian=True) -> theta, U, resnorm, Z (sorted by |theta|)."""
    n = H.shape[1]
    theta, U = eigh(H[:n, :n])
    order = np.argsort(np.abs(theta))
    theta, U = theta[order], U[:, order]
    resnorm = np.abs(H[n, n - 1] * U[n - 1, :]) if H.shape[0] > n else np.zeros(n)
    Z = None
    if V is not None:
        if isinstance(V, torch.Tensor):
            Z = torch.matmul(V[:, :n], dv.to_dev_f64(U))
        else:
            Z = dv.to_host(torch.matmul(dv.to_dev_f64(np.ascontiguousarray(V[:, :n])), dv.to_dev_f64(U)))
    return theta, U, resnorm, Z


def find_ritz_eigenvalues(h, v, threshold=1.e-2, eigenvalues=False, filename=None):
    """interfaces/deflationlib.py:204-219 -> ``(Z, r)`` (and the selected theta if asked)."""
    eig, u, resnorm, z = krypy_ritz(h, V=v, hermitian=True)
    sel = eig < threshold
    r = int(np.count_nonzero(sel))
    if eigenvalues:
        idx = np.nonzero(sel)[0]
        zsel = z[:, torch.as_tensor(idx, device=z.device)] if isinstance(z, torch.Tensor) else z[:, idx]
        return zsel, r, eig[sel]
    return z[:, :r], r
