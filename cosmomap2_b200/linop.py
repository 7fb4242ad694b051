"""
Operator algebra with the surface of the third-party ``linop`` package the reference builds on
(interfaces/linearoperators.py:14 ``import linop.linop as lp``; interfaces/blkop.py:1-2), so the
reference's call sites -- ``P.T*N*P``, ``Mbd*R + Zd*E*Zd.T``, ``lp.IdentityOperator(n)``,
``op*ndarray``, ``.T .H .shape .dtype .symmetric .matvec .to_array()`` -- work unchanged.

Differences from a CPU linop, all invisible to those call sites:
  * vectors live in HBM while they flow through a composed operator: ``op*ndarray`` uploads once,
    every factor consumes/produces a CUDA fp64 tensor (``_apply``), and one result is downloaded.
    ``op*cuda_tensor`` stays on the device end to end (what the device PCG uses);
  * products keep their factor list, so a known pattern (``P.T*N*P`` ...) is replaced by a fused
    kernel at first use (``register_fuser``).
"""
import logging

import numpy as np
import torch

from . import _device as dv

null_log = logging.getLogger("cosmomap2_b200.linop")
null_log.addHandler(logging.NullHandler())
null_log.propagate = False


class ShapeError(Exception):
    """Raised when a vector or operator of the wrong shape is used (linop.ShapeError)."""


class BaseLinearOperator(object):
    def __init__(self, nargin, nargout, symmetric=False, hermitian=False, dtype=np.float64, **kwargs):
        self.__nargin = int(nargin)
        self.__nargout = int(nargout)
        self.__symmetric = bool(symmetric)
        self.__hermitian = bool(hermitian)
        self.__shape = (self.__nargout, self.__nargin)
        self.__dtype = np.dtype(dtype)
        self._nMatvec = 0
        self.logger = kwargs.get("logger", null_log)

    nargin = property(lambda self: self.__nargin)
    nargout = property(lambda self: self.__nargout)
    symmetric = property(lambda self: self.__symmetric)
    hermitian = property(lambda self: self.__hermitian)
    shape = property(lambda self: self.__shape)
    dtype = property(lambda self: self.__dtype)
    nMatvec = property(lambda self: self._nMatvec)


class LinearOperator(BaseLinearOperator):
    """``LinearOperator(nargin, nargout, matvec, rmatvec=None, symmetric=False, ...)``.

    ``device=True`` marks ``matvec``/``rmatvec`` as device-native (CUDA fp64 tensor in and out);
    operators built by user code from NumPy callables (``device=False``, the default) are wrapped
    so they can still take part in device-resident compositions.
    """

    def __init__(self, nargin, nargout, matvec, rmatvec=None, **kwargs):
        device = kwargs.pop("device", False)
        adjoint_of = kwargs.pop("adjoint_of", None)
        super(LinearOperator, self).__init__(nargin, nargout, **kwargs)
        self._device_native = bool(device)
        self._mv = matvec
        self._rmv = rmatvec
        self._adjoint_of = adjoint_of
        self.__H = None

    # -- transpose ------------------------------------------------------------------------------
    def _make_transpose(self):
        if self._rmv is None:
            return None
        return LinearOperator(self.nargout, self.nargin, matvec=self._rmv, rmatvec=self._mv,
                              device=self._device_native, adjoint_of=self, dtype=self.dtype)

    @property
    def T(self):
        if self.symmetric:
            return self
        if self._adjoint_of is not None:
            return self._adjoint_of
        if self.__H is None:
            self.__H = self._make_transpose()
        return self.__H

    H = T

    # -- application ----------------------------------------------------------------------------
    def _apply(self, x):
        """Device-level apply: CUDA fp64 tensor -> CUDA fp64 tensor."""
        self._nMatvec += 1
        if self._device_native:
            return self._mv(x)
        return dv.to_dev_f64(np.asarray(self._mv(dv.to_host(x)), dtype=np.float64))

    def matvec(self, x):
        if isinstance(x, torch.Tensor):
            if x.shape[0] != self.nargin:
                raise ShapeError("Multiplying with vector of wrong shape.")
            return self._apply(dv.to_dev_f64(x))
        x = np.asanyarray(x)
        col = x.ndim == 2 and x.shape[1] == 1
        if col:
            x = x[:, 0]
        if x.ndim != 1 or x.shape[0] != self.nargin:
            raise ShapeError("Multiplying with vector of wrong shape.")
        if not self._device_native:
            self._nMatvec += 1
            y = np.asarray(self._mv(np.asarray(x, dtype=np.float64)))
        else:
            y = dv.to_host(self._apply(dv.to_dev_f64(x)))
        return y.reshape(-1, 1) if col else y

    def rmatvec(self, x):
        t = self.T
        if t is None:
            raise NotImplementedError("operator has no transpose")
        return t.matvec(x)

    def to_array(self):
        n, m = self.shape
        out = np.empty((n, m), dtype=np.float64)
        e = np.zeros(m)
        for j in range(m):
            e[j] = 1.0
            out[:, j] = self.matvec(e)
            e[j] = 0.0
        return out

    def __call__(self, x):
        return self.__mul__(x)

    def dot(self, x):
        return self.__mul__(x)

    # -- algebra --------------------------------------------------------------------------------
    def __mul__(self, x):
        if np.isscalar(x):
            return _ScaledLO(self, x)
        if isinstance(x, BaseLinearOperator):
            if self.nargin != x.nargout:
                raise ShapeError("Cannot multiply operators together")
            return _ProductLO(_factors(self) + _factors(x))
        if isinstance(x, (np.ndarray, list, tuple, torch.Tensor)):
            return self.matvec(np.asarray(x) if isinstance(x, (list, tuple)) else x)
        raise ValueError("Cannot multiply")

    def __rmul__(self, x):
        if np.isscalar(x):
            return _ScaledLO(self, x)
        raise ValueError("Cannot multiply")

    def __add__(self, other):
        return _SumLO(self, other, 1.0)

    def __sub__(self, other):
        return _SumLO(self, other, -1.0)

    def __neg__(self):
        return _ScaledLO(self, -1.0)

    def __truediv__(self, a):
        if np.isscalar(a):
            return _ScaledLO(self, 1.0 / a)
        raise ValueError("Cannot divide")


def _factors(op):
    return list(op.factors) if isinstance(op, _ProductLO) else [op]


_fusers = []


def register_fuser(fn):
    """``fn(factors) -> new factor list or None``: lets linearoperators.py replace known
    sub-chains (e.g. [P.T, N, P]) by one fused device operator."""
    _fusers.append(fn)
    return fn


class _ProductLO(LinearOperator):
    def __init__(self, factors):
        self.factors = list(factors)
        self._plan = None
        super(_ProductLO, self).__init__(self.factors[-1].nargin, self.factors[0].nargout,
                                         matvec=self._run, device=True,
                                         dtype=np.result_type(*[f.dtype for f in self.factors]))

    def _make_transpose(self):
        ts = [f.T for f in reversed(self.factors)]
        if any(t is None for t in ts):
            return None
        t = _ProductLO(ts)
        t._adjoint_of = self
        return t

    @property
    def T(self):
        if self._adjoint_of is not None:
            return self._adjoint_of
        if getattr(self, "_T", None) is None:
            self._T = self._make_transpose()
        return self._T

    H = T

    def planned(self):
        if self._plan is None:
            plan = self.factors
            changed = True
            while changed:
                changed = False
                for fuse in _fusers:
                    new = fuse(plan)
                    if new is not None:
                        plan = new
                        changed = True
            self._plan = plan
        return self._plan

    def _run(self, x):
        for f in reversed(self.planned()):
            x = f._apply(x)
        return x


_sum_fusers = []


def register_sum_fuser(fn):
    """``fn(sum_op) -> replacement operator or None`` (e.g. the two-level preconditioner)."""
    _sum_fusers.append(fn)
    return fn


class _SumLO(LinearOperator):
    def __init__(self, a, b, sign):
        if not isinstance(b, BaseLinearOperator):
            raise ValueError("Cannot add")
        if a.shape != b.shape:
            raise ShapeError("Cannot add")
        self.a, self.b, self.sign = a, b, float(sign)
        self._fused = False
        super(_SumLO, self).__init__(a.nargin, a.nargout, matvec=self._run, device=True,
                                     symmetric=a.symmetric and b.symmetric,
                                     dtype=np.result_type(a.dtype, b.dtype))

    def _make_transpose(self):
        if self.a.T is None or self.b.T is None:
            return None
        t = _SumLO(self.a.T, self.b.T, self.sign)
        t._adjoint_of = self
        return t

    def _run(self, x):
        if self._fused is False:
            self._fused = None
            for fuse in _sum_fusers:
                rep = fuse(self)
                if rep is not None:
                    self._fused = rep
                    break
        if self._fused is not None:
            return self._fused._apply(x)
        ya = self.a._apply(x)
        yb = self.b._apply(x)
        if ya is x or ya.data_ptr() == x.data_ptr():      # identity returns its input: do not clobber it
            ya = ya.clone()
        dv.call("cm2_axpby", self.sign, dv.ptr(yb), 1.0, dv.ptr(ya), ya.numel(), dv.stream())
        return ya


class _ScaledLO(LinearOperator):
    def __init__(self, op, a):
        self.op, self.a = op, float(a)
        super(_ScaledLO, self).__init__(op.nargin, op.nargout, matvec=self._run, device=True,
                                        symmetric=op.symmetric, dtype=op.dtype)

    def _make_transpose(self):
        if self.op.T is None:
            return None
        t = _ScaledLO(self.op.T, self.a)
        t._adjoint_of = self
        return t

    def _run(self, x):
        y = self.op._apply(x)
        if y is x or y.data_ptr() == x.data_ptr():
            y = y.clone()
        dv.call("cm2_axpby", 0.0, dv.ptr(y), self.a, dv.ptr(y), y.numel(), dv.stream())
        return y


class IdentityOperator(LinearOperator):
    def __init__(self, nargin, **kwargs):
        kwargs.pop("symmetric", None)
        super(IdentityOperator, self).__init__(nargin, nargin, symmetric=True, matvec=lambda x: x,
                                               device=True, **kwargs)


class DiagonalOperator(LinearOperator):
    def __init__(self, diag, **kwargs):
        d = np.asarray(dv.to_host(diag) if isinstance(diag, torch.Tensor) else diag, dtype=np.float64)
        if d.ndim != 1:
            raise ValueError("Input must be 1-d array")
        self.__diag = d.copy()
        self._diag_dev = None
        kwargs.pop("symmetric", None)
        super(DiagonalOperator, self).__init__(d.shape[0], d.shape[0], symmetric=True, matvec=self._run,
                                               device=True, **kwargs)

    @property
    def diag(self):
        return self.__diag

    def _run(self, x):
        if self._diag_dev is None:
            self._diag_dev = dv.to_dev_f64(self.__diag)
        # per-sample blocks of size 1: the white-noise kernel with blk_start = NULL, blocksize = 1
        out = torch.empty_like(x)
        dv.call("cm2_noise_white_apply", dv.ptr(self._diag_dev), x.numel(), 1, None, dv.ptr(x), dv.ptr(out),
                x.numel(), dv.stream())
        return out


class ZeroOperator(LinearOperator):
    def __init__(self, nargin, nargout, **kwargs):
        super(ZeroOperator, self).__init__(nargin, nargout, matvec=lambda x: dv.zeros_f64(nargout),
                                           rmatvec=lambda x: dv.zeros_f64(nargin), device=True, **kwargs)
