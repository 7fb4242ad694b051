"""
Device-resident preconditioned conjugate gradient with the call signature, recurrence, stopping
rule and return values of ``scipy.sparse.linalg.cg`` (SciPy 1.18, _isolve/iterative.py:383-431),
which is what the reference calls (src/test_BD_precond_onto_real_data.py:47,
src/test_M2_precond_onto_real_data.py:117, tests/test_2level_preconditioner.py:52).

    r = b - A x0                     (b itself when x0 == 0)
    each iteration:  FIRST test ||r||_2 < max(atol, rtol*||b||_2) -> return (x, 0)
                     z = M r ; rho = r.z ; p = z + (rho/rho_prev) p ; q = A p
                     alpha = rho / p.q ; x += alpha p ; r -= alpha q ; callback(x)
    after maxiter iterations -> return (x, maxiter)

All vectors stay in HBM; alpha/beta/rho live in a device workspace (cm2_pcg_update_*), and the
only per-iteration host read is ||r||^2 (8 bytes) for the exit test, so the iteration count is
exactly SciPy's.  NumPy ``b`` in -> NumPy ``x`` out; CUDA tensor in -> CUDA tensor out.
"""
import numpy as np
import torch

from . import _device as dv
from . import linop as lp


class _Identity(object):
    def _apply(self, x):
        return x


def _as_device_operator(op, n):
    if op is None:
        return _Identity()
    if isinstance(op, lp.LinearOperator):
        return op
    if hasattr(op, "matvec") and hasattr(op, "shape"):
        return lp.LinearOperator(op.shape[1], op.shape[0], matvec=op.matvec, device=False)
    arr = np.asarray(op, dtype=np.float64)
    return lp.LinearOperator(arr.shape[1], arr.shape[0], matvec=lambda v: arr.dot(v), device=False)


class PCG(object):
    """The PCG state machine: ``start(b, x0)`` then ``step()`` per iteration.

    ``rnorm`` is ||r||_2 as SciPy tests it at the top of the next iteration.  ``a_events`` (a list),
    when set, receives a (start, stop) CUDA-event pair around every A apply -- bench.py uses it
    to time the dominant kernel inside the timed region.
    """

    def __init__(self, A, M, n):
        dv.require_cuda()
        self.n = int(n)
        self.A = _as_device_operator(A, n)
        self.M = _as_device_operator(M, n)
        self.scal = dv.zeros_f64(8)
        self.x = dv.zeros_f64(n)
        self.r = dv.zeros_f64(n)
        self.p = dv.empty_f64(n)
        self.iteration = 0
        self.rnorm = 0.0
        self.a_events = None

    def _dot(self, a, b, slot):
        dv.call("cm2_dot", dv.ptr(a), dv.ptr(b), self.n, dv.ptr(self.scal) + 8 * slot, dv.stream())

    def norm(self, v):
        self._dot(v, v, 6)
        return float(np.sqrt(self.scal[6].item()))

    def start(self, b, x0=None, need_norm=True):
        """x <- x0 (or 0), r <- b - A x0.  ``b`` must be a CUDA fp64 tensor."""
        n = self.n
        self.iteration = 0
        self.r.copy_(b)
        if x0 is None:
            self.x.zero_()
        else:
            self.x.copy_(x0)
            if bool(torch.any(self.x != 0).item()):
                ax = self.A._apply(self.x)
                dv.call("cm2_axpby", -1.0, dv.ptr(ax), 1.0, dv.ptr(self.r), n, dv.stream())
        if need_norm:
            self.rnorm = self.norm(self.r)

    def step(self, read_norm=True):
        n, st = self.n, dv.stream
        z = self.M._apply(self.r)
        dv.call("cm2_pcg_update_p", dv.ptr(self.r), dv.ptr(z), dv.ptr(self.p), n, dv.ptr(self.scal),
                int(self.iteration == 0), st())
        if self.a_events is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            q = self.A._apply(self.p)
            e1.record()
            self.a_events.append((e0, e1))
        else:
            q = self.A._apply(self.p)
        dv.call("cm2_pcg_update_xr", dv.ptr(self.p), dv.ptr(q), dv.ptr(self.x), dv.ptr(self.r), n,
                dv.ptr(self.scal), st())
        self.iteration += 1
        if read_norm:
            self.rnorm = float(np.sqrt(self.scal[3].item()))     # the one 8-byte host read per iteration
        return self.rnorm


def cg(A, b, x0=None, *, rtol=1e-5, atol=0., maxiter=None, M=None, callback=None, tol=None,
       residuals=None):
    """``x, info = cg(A, b, x0=None, rtol=1e-5, atol=0., maxiter=None, M=None, callback=None)``.

    ``tol`` is the legacy name of ``rtol`` used by the reference's call sites.  ``residuals``, if a
    list, receives ||r||_2 as tested at the top of every iteration (no extra device work).
    """
    dv.require_cuda()
    if tol is not None:
        rtol = tol
    if atol is None or atol < 0:
        raise ValueError("'cg' called with invalid `atol`=%r; if set, `atol` must be a real, "
                         "non-negative number." % (atol,))
    want_numpy = not isinstance(b, torch.Tensor)
    b_in = np.asarray(b, dtype=np.float64).reshape(-1) if want_numpy else b.reshape(-1)
    n = b_in.shape[0]
    if isinstance(A, lp.LinearOperator) and A.shape != (n, n):
        raise ValueError("A and b have incompatible dimensions")
    solver = PCG(A, M, n)
    bd = dv.to_dev_f64(b_in)
    bnrm2 = solver.norm(bd)
    atol = max(float(atol), float(rtol) * bnrm2)
    if bnrm2 == 0:
        return (b_in.copy() if want_numpy else bd.clone()), 0
    if maxiter is None:
        maxiter = n * 10
    x0d = None
    if x0 is not None:
        x0d = dv.to_dev_f64(x0).reshape(-1)
        if x0d.shape[0] != n:
            raise ValueError("shapes of A and x0 are incompatible")
    solver.start(bd, x0d)

    def out(v):
        return dv.to_host(v) if want_numpy else v

    for _ in range(maxiter):
        if residuals is not None:
            residuals.append(solver.rnorm)
        if solver.rnorm < atol:
            return out(solver.x), 0
        solver.step()
        if callback:
            callback(out(solver.x))
    return out(solver.x), maxiter
