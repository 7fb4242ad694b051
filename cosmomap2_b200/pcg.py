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

All vectors and scalars stay in HBM (16-double device workspace, see include/cosmomap2_b200.h).
SciPy's exit test is evaluated ON THE DEVICE by the kernel that updates r; it raises a `done` flag
that turns every later update kernel into a no-op.  The host therefore queues iteration k+1 while
iteration k is still running and reads the flag one iteration late (8 bytes, pinned, async): the
GPU never idles, yet x, the iteration count and info are exactly SciPy's.
When M is the block-diagonal preconditioner its apply is pixel-local and is folded into the
kernel that updates r, together with the next search direction: an iteration is the A apply plus
ONE cooperative launch (cm2_pcg_bd_iter).
NumPy ``b`` in -> NumPy ``x`` out; CUDA tensor in -> CUDA tensor out.
"""
import numpy as np
import torch

from . import _device as dv
from . import linop as lp

NSCAL = 16


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
    """The PCG state machine: ``start(b, x0, atol)``, then ``step_async()`` per iteration and
    ``state()`` / ``poll()`` to read the device scalars."""

    def __init__(self, A, M, n):
        dv.require_cuda()
        from .linearoperators import BlockDiagonalPreconditionerLO
        self.n = int(n)
        self.A = _as_device_operator(A, n)
        self.M = _as_device_operator(M, n)
        self.bd = self.M if isinstance(self.M, BlockDiagonalPreconditionerLO) and self.M.size == self.n else None
        self.scal = dv.zeros_f64(NSCAL)
        self.x = dv.zeros_f64(n)
        self.r = dv.zeros_f64(n)
        self.p = dv.empty_f64(n)
        self.z = dv.empty_f64(n) if self.bd is not None else None
        self._pin = [torch.empty(NSCAL, dtype=torch.float64).pin_memory() for _ in range(2)]
        self._ev = [torch.cuda.Event(), torch.cuda.Event()]
        # one cooperative launch for the pixel-domain tail; if the device refuses cooperative launches
        # (attribute 0 under some sharing modes, or a refused launch at run time) the 3-kernel tail runs
        self._coop = True
        self._queued = 0          # iterations launched since start()
        self._snap = 0            # snapshots enqueued

    # ---- scalars -----------------------------------------------------------------------------
    def norm(self, v):
        out = self.scal[15:16]
        dv.call("cm2_dot", dv.ptr(v), dv.ptr(v), self.n, dv.ptr(out), dv.stream())
        return float(np.sqrt(out.item()))

    def state(self):
        """Synchronous read of (rnorm, done, iterations) from the device workspace."""
        s = self.scal.cpu()
        return float(np.sqrt(s[3].item())), bool(s[7].item() != 0.0), int(s[8].item())

    def _snapshot(self):
        k = self._snap % 2
        self._pin[k].copy_(self.scal, non_blocking=True)
        self._ev[k].record()
        self._snap += 1

    def _read_snapshot(self, idx):
        k = idx % 2
        self._ev[k].synchronize()
        s = self._pin[k]
        return float(np.sqrt(s[3].item())), bool(s[7].item() != 0.0), int(s[8].item())

    # ---- recurrence --------------------------------------------------------------------------
    def start(self, b, x0=None, atol=0.0, rtol=0.0):
        """x <- x0 (or 0), r <- b - A x0, device scalars reset.  ``b`` is a CUDA fp64 tensor.
        ``rtol`` > 0 folds SciPy's ``atol = max(atol, rtol*||b||)`` in (one synchronising norm)."""
        n, st = self.n, dv.stream
        if x0 is None and self.bd is not None:
            # x0 = 0 with M_BD: r = b, x = 0, z = M r, rho, ||r||^2 and atol = max(atol, rtol ||b||) in ONE pass
            dv.call("cm2_pcg_bd_reset", dv.ptr(self.bd._inv_dev), self.bd._n, self.bd.pol, dv.ptr(self.r),
                    dv.ptr(self.z), dv.ptr(self.scal), float(atol), float(rtol), dv.ptr(b), dv.ptr(self.x), dv.ptr(self.p), st())
            self._queued = 0
            return
        if rtol:
            atol = max(float(atol), float(rtol) * self.norm(b))
        self.r.copy_(b)
        if x0 is None:
            self.x.zero_()
        else:
            self.x.copy_(x0)
            if bool(torch.any(self.x != 0).item()):
                ax = self.A._apply(self.x)
                dv.call("cm2_axpby", -1.0, dv.ptr(ax), 1.0, dv.ptr(self.r), n, st())
        if self.bd is not None:
            dv.call("cm2_pcg_bd_reset", dv.ptr(self.bd._inv_dev), self.bd._n, self.bd.pol, dv.ptr(self.r),
                    dv.ptr(self.z), dv.ptr(self.scal), float(atol), 0.0, None, None, dv.ptr(self.p), st())
        else:
            dv.call("cm2_pcg_reset", dv.ptr(self.r), n, dv.ptr(self.scal), float(atol), st())
        self._queued = 0

    def _apply_A(self, p):
        # q = A p is consumed inside the iteration: a transient (buffer-aliasing) result is fine
        f = getattr(self.A, "apply_transient", None)
        return f(p) if f is not None else self.A._apply(p)

    def step_async(self):
        """Queue one iteration (no host synchronisation)."""
        n, st = self.n, dv.stream
        if self.bd is not None:
            # p already holds this iteration's search direction (bd_reset / the previous bd_iter)
            q = self._apply_A(self.p)
            if self._coop:
                try:
                    dv.call("cm2_pcg_bd_iter", dv.ptr(self.bd._inv_dev), self.bd._n, self.bd.pol, dv.ptr(self.p),
                            dv.ptr(q), dv.ptr(self.x), dv.ptr(self.r), dv.ptr(self.z), dv.ptr(self.scal), st())
                except dv._cabi.Cm2Error:
                    self._coop = False          # cooperative launch refused (e.g. GPU shared): 3-kernel tail
            if not self._coop:
                dv.call("cm2_pcg_bd_update", dv.ptr(self.bd._inv_dev), self.bd._n, self.bd.pol, dv.ptr(self.p),
                        dv.ptr(q), dv.ptr(self.x), dv.ptr(self.r), dv.ptr(self.z), dv.ptr(self.scal), st())
                dv.call("cm2_pcg_bd_update_p", dv.ptr(self.z), dv.ptr(self.p), n, dv.ptr(self.scal), st())
        else:
            z = self.M._apply(self.r)
            dv.call("cm2_pcg_update_p", dv.ptr(self.r), dv.ptr(z), dv.ptr(self.p), n, dv.ptr(self.scal), st())
            q = self._apply_A(self.p)
            dv.call("cm2_pcg_update_xr", dv.ptr(self.p), dv.ptr(q), dv.ptr(self.x), dv.ptr(self.r), n,
                    dv.ptr(self.scal), st())
        self._queued += 1

    def tick(self):
        """Enqueue an async snapshot of the device scalars and return the PREVIOUS snapshot
        (rnorm, done, iterations) -- the one-iteration-late read the solve loop uses."""
        self._snapshot()
        if self._snap >= 2:
            return self._read_snapshot(self._snap - 2)
        return None

    def step(self):
        """One iteration, synchronous: returns ||r||_2 after it."""
        self.step_async()
        return self.state()[0]


def _run_loop(solver, maxiter, residuals):
    """The asynchronous solve loop: the flag of iteration k is read while iteration k+1 is already
    queued.  Returns info (0 = converged, maxiter = not)."""
    s0 = solver._snap
    solver._snapshot()                      # snapshot s0: state before iteration 0
    for it in range(maxiter):
        solver.step_async()                 # no-op on the device if `done` was already raised
        solver._snapshot()                  # snapshot s0+it+1: state after iteration `it`
        rnorm, done, _iters = solver._read_snapshot(s0 + it)   # state at the TOP of iteration `it`
        if residuals is not None:
            residuals.append(rnorm)
        if done:
            return 0                        # the queued iteration `it` did nothing
    return maxiter


def _cg_sharded(solver, A, b_in, want_numpy, rtol, atol, maxiter, residuals, gather):
    """cg() over the pixel-sharded multi-GPU solver (distributed.ShardedPCG).  Returns None if the
    peer-memory exchange failed on ANY rank (every rank then takes the replicated NCCL path)."""
    lo, hi = solver.elo, solver.ehi
    if want_numpy:
        bs = dv.to_dev_f64(b_in[lo:hi])     # only this rank's slice crosses PCIe
    else:
        bs = b_in[lo:hi]
    solver.start(bs, None, atol, rtol)
    res = [] if residuals is not None else None
    info = _run_loop(solver, maxiter, res)
    from . import distributed
    if distributed.agree_failed(solver.failed(), solver.group):
        A.disable_p2p("a peer-flag wait of the sharded PCG timed out")
        return None
    if residuals is not None:
        residuals.extend(res)
    if gather == "shard":
        x = solver.x
        return (dv.to_host(x) if want_numpy else x.clone()), info
    x = solver.gather_x()
    return (dv.to_host(x) if want_numpy else x), info


def make_solver(A, M, n):
    """The PCG state machine for (A, M): the pixel-sharded multi-GPU solver when A is a
    ``distributed.AllReduceLO`` over peer memory and M the block-diagonal preconditioner, else ``PCG``."""
    f = getattr(A, "sharded_solver", None)
    s = f(M) if f is not None else None
    return s if s is not None else PCG(A, M, n)


def cg(A, b, x0=None, *, rtol=1e-5, atol=0., maxiter=None, M=None, callback=None, tol=None,
       residuals=None, gather="all"):
    """``x, info = cg(A, b, x0=None, rtol=1e-5, atol=0., maxiter=None, M=None, callback=None)``.

    ``tol`` is the legacy name of ``rtol`` used by the reference's call sites.  ``residuals``, if a
    list, receives ||r||_2 as tested at the top of every iteration.  Multi-GPU (A a
    ``distributed.AllReduceLO``): ``gather="all"`` returns the full x on every rank, ``"shard"`` only
    this rank's pixel slice ``x[pol*lo:pol*hi]`` (``distributed.partition_pixels``) when the
    pixel-sharded solver runs, so that b and x cross PCIe once per job instead of once per rank.
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
    if maxiter is None:
        maxiter = n * 10
    if x0 is None and callback is None:
        f = getattr(A, "sharded_solver", None)
        sharded = f(M) if f is not None else None
        if sharded is not None:
            out_ = _cg_sharded(sharded, A, b_in, want_numpy, rtol, atol, maxiter, residuals, gather)
            if out_ is not None:
                return out_
    while True:
        x, info = _cg_replicated(A, b_in, x0, want_numpy, rtol, atol, maxiter, M, callback, residuals)
        rec = getattr(A, "recover", None)
        if rec is None or not rec():
            return x, info
        if residuals is not None:           # the peer-memory all-reduce failed mid-solve on some rank:
            del residuals[:]                # every rank has switched to NCCL; solve again


def _cg_replicated(A, b_in, x0, want_numpy, rtol, atol, maxiter, M, callback, residuals):
    n = b_in.shape[0]
    # the state machine (five n-vectors, scalars, pinned snapshot buffers) is kept on the operator between calls with
    # the same preconditioner: start() resets all of it, results are handed out as copies
    solver = None
    cached = getattr(A, "_cm2_pcg", None) if isinstance(A, lp.LinearOperator) else None
    if cached is not None and cached[0] is M and cached[1].n == n:
        solver = cached[1]
    if solver is None:
        solver = PCG(A, M, n)
        if isinstance(A, lp.LinearOperator) and (M is None or isinstance(M, lp.LinearOperator)):
            try:
                A._cm2_pcg = (M, solver)
            except AttributeError:
                pass
    bd = dv.to_dev_f64(b_in)
    x0d = None
    if x0 is not None:
        x0d = dv.to_dev_f64(x0).reshape(-1)
        if x0d.shape[0] != n:
            raise ValueError("shapes of A and x0 are incompatible")
    if solver.bd is not None and x0d is None:
        # M_BD, x0 = 0: SciPy's atol = max(atol, rtol ||b||) and its b = 0 exit (x = 0, info = 0) are evaluated by
        # the start kernel on the device -- no host round trip for ||b||
        solver.start(bd, None, float(atol), float(rtol))
    else:
        bnrm2 = solver.norm(bd)
        atol = max(float(atol), float(rtol) * bnrm2)
        if bnrm2 == 0:
            return (b_in.copy() if want_numpy else bd.clone()), 0
        solver.start(bd, x0d, atol)

    def out(v):
        return dv.to_host(v) if want_numpy else v.clone()

    if callback is not None:
        # the callback wants x after every iteration: synchronous loop
        rnorm, done, _ = solver.state()
        for _ in range(maxiter):
            if residuals is not None:
                residuals.append(rnorm)
            if done:
                return out(solver.x), 0
            solver.step_async()
            rnorm, done, _ = solver.state()
            callback(out(solver.x))
        return out(solver.x), maxiter

    info = _run_loop(solver, maxiter, residuals)
    return out(solver.x), info
