"""
ORACLE (test infrastructure, not product code): Krylov / deflation helpers.

* ``arnoldi``, ``build_hess``, ``build_Z`` restate interfaces/deflationlib.py:17-184 (in-tree
  Euclidean MGS Arnoldi).  ``build_Z`` in the reference picks ROWS ``y[i]`` of the ``eigh``
  eigenvector matrix and calls ``w.T`` on a Python list (deflationlib.py:172-174, 183) -- it is
  only reachable from an uncalled function (src/test_M2_precond_onto_real_data.py:13-50).  Here
  it is restated with eigenvector COLUMNS; this is a documented deviation.
* ``run_krypy_arnoldi`` / ``find_ritz_eigenvalues`` (deflationlib.py:187-219) delegate to the
  third-party ``krypy`` (unpinned, .travis.yml:37; absent from /root/reference and from this
  image).  PARITY UNPINNED: restated from krypy's published algorithm (krypy.utils.Arnoldi with a
  preconditioner ``M``: V = M P, <P_i, V_j> = delta_ij, modified Gram-Schmidt;
  krypy.utils.ritz(H, V, hermitian=True): eigh of the square Hessenberg, Ritz vectors V U, sorted
  by |theta|).  Pinned only by identities: M A V_m = V_{m+1} H, V^T P = I, and the Ritz values
  against dense eigh of M^{1/2} A M^{1/2} on small problems (tests/test_oracle_krylov.py).
  ``ortho='dmgs'`` (two passes) is offered because single-pass MGS loses bi-orthogonality
  (SURVEY.md appendix A.8); the reference's call leaves krypy's default 'mgs'.
"""
import numpy as np
from scipy.linalg import eigh

from .operators import dgemm, norm2


def arnoldi(A, b, x0=None, tol=1e-5, maxiter=1000, inner_m=30):
    """interfaces/deflationlib.py:17-113, literally (including the stop test at :101)."""
    if not np.isfinite(b).all():
        raise ValueError("RHS must contain only finite numbers")
    matvec = A.matvec
    b_norm = norm2(b)
    if b_norm == 0:
        b_norm = 1
    r_outer = b - matvec(x0)
    r_norm = norm2(r_outer)
    if r_norm < tol * b_norm or r_norm < tol:
        return None, None, 0
    vs = [r_outer / r_norm]
    hs = []
    for j in range(1, 1 + inner_m):
        v_new = matvec(vs[j - 1])
        v_new2 = v_new.copy()
        hcur = []
        for v in vs:
            alpha = np.dot(v, v_new)
            hcur.append(alpha)
            v_new2 = v_new2 - alpha * v     # axpy(v, v_new2, n, -alpha): in place on v_new2
            v_new = v_new2
        hcur.append(norm2(v_new))
        v_new = v_new / hcur[-1]
        if abs(v_new[j] * hcur[-1]) <= tol:
            hs.append(hcur)
            return vs, hs, j
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
    """interfaces/deflationlib.py:140-184 with eigenvector columns (see module docstring)."""
    m = len(z)
    sel = [i for i in range(m) if abs(z[i]) <= eps]
    r = len(sel)
    if r == 0:
        raise RuntimeError("No Ritz eigenvalue are found smaller than fixed threshold %.1g " % eps)
    W = np.asarray(w)[:m].T                  # npix x m
    Z = W.dot(np.asarray(y)[:, sel])
    return Z, r


def krypy_arnoldi(A, x0, M=None, maxiter=None, ortho="mgs", tol_invariant=1e-14):
    """krypy.utils.arnoldi(A, v, M=M, maxiter, ortho) -> V, H, P (see module docstring)."""
    N = len(x0)
    maxiter = N if maxiter is None else min(maxiter, N)
    p = np.array(x0, dtype=np.float64).reshape(N)
    v = M.matvec(p) if M is not None else p
    nrm = np.sqrt(np.dot(p, v))
    V = np.zeros((N, maxiter + 1))
    P = np.zeros((N, maxiter + 1)) if M is not None else V
    H = np.zeros((maxiter + 1, maxiter))
    V[:, 0] = v / nrm
    if M is not None:
        P[:, 0] = p / nrm
    k = 0
    invariant = False
    while k < maxiter and not invariant:
        w = A.matvec(V[:, k])
        for _ in range(2 if ortho == "dmgs" else 1):
            for j in range(k + 1):
                a = np.dot(V[:, j], w)
                H[j, k] += a
                w = w - a * P[:, j]
        Mw = M.matvec(w) if M is not None else w
        hk = np.sqrt(abs(np.dot(w, Mw)))
        H[k + 1, k] = hk
        if hk / np.abs(H[:k + 2, :k + 1]).max() <= tol_invariant:
            invariant = True
        else:
            if M is not None:
                P[:, k + 1] = w / hk
            V[:, k + 1] = Mw / hk
        k += 1
    if invariant:
        return V[:, :k], H[:k, :k], P[:, :k]
    return V[:, :k + 1], H[:k + 1, :k], P[:, :k + 1]


def run_krypy_arnoldi(A, x0, M, tol, maxiter=None, ortho="mgs"):
    """interfaces/deflationlib.py:187-202."""
    v, h, p = krypy_arnoldi(A, x0, M=M, maxiter=maxiter, ortho=ortho)
    return v, h, v.shape[1]


def krypy_ritz(H, V=None, hermitian=True):
    """krypy.utils.ritz(H, V=V, hermitian=True) -> theta, U, resnorm, Z (sorted by |theta|)."""
    n = H.shape[1]
    Hs = H[:n, :n]
    theta, U = eigh(Hs)
    order = np.argsort(np.abs(theta))
    theta, U = theta[order], U[:, order]
    if H.shape[0] > n:
        resnorm = np.abs(H[n, n - 1] * U[n - 1, :])
    else:
        resnorm = np.zeros(n)
    Z = None if V is None else V[:, :n].dot(U)
    return theta, U, resnorm, Z


def find_ritz_eigenvalues(h, v, threshold=1.e-2, eigenvalues=False):
    """interfaces/deflationlib.py:204-219."""
    eig, u, resnorm, z = krypy_ritz(h, V=v, hermitian=True)
    sel = eig < threshold
    r = int(np.count_nonzero(sel))
    if eigenvalues:
        return z[:, sel], r, eig[sel]
    return z[:, :r], r
