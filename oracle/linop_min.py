"""
ORACLE (test infrastructure, not product code) -- minimal stand-in for the third-party
``linop`` package the reference builds every operator on (interfaces/linearoperators.py:14,
interfaces/blkop.py:1-2).  ``linop`` is not vendored in /root/reference and is unpinned
(.travis.yml:33-39); it contains no arithmetic of its own beyond composition, so this file
restates only the operator algebra the reference actually uses:
``op*ndarray``, ``op*op``, ``op+op``, ``op-op``, ``-op``, ``scalar*op``, ``.T``, ``.H``,
``.shape``, ``.dtype``, ``.symmetric``, ``.matvec``, ``.to_array()``, ``IdentityOperator``,
``DiagonalOperator``, ``ShapeError``, ``null_log``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything under oracle/.
"""
import logging

import numpy as np

null_log = logging.getLogger("oracle.linop")
null_log.addHandler(logging.NullHandler())
null_log.propagate = False


class ShapeError(Exception):
    """Raised when a vector or operator of the wrong shape is used (linop.ShapeError)."""


class BaseLinearOperator(object):
    def __init__(self, nargin, nargout, symmetric=False, hermitian=False,
                 dtype=np.float64, **kwargs):
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
    """``LinearOperator(nargin, nargout, matvec, rmatvec=None, symmetric=False, ...)``."""

    def __init__(self, nargin, nargout, matvec, rmatvec=None, **kwargs):
        super(LinearOperator, self).__init__(nargin, nargout, **kwargs)
        adjoint_of = kwargs.get("adjoint_of", None)
        self.__matvec = matvec
        if self.symmetric:
            self.__H = self
        elif adjoint_of is not None:
            self.__H = adjoint_of
        elif rmatvec is not None:
            self.__H = LinearOperator(nargout, nargin, matvec=rmatvec, rmatvec=matvec,
                                      adjoint_of=self, dtype=self.dtype)
        else:
            self.__H = None

    @property
    def T(self):
        return self.__H

    @property
    def H(self):
        return self.__H

    def matvec(self, x):
        self._nMatvec += 1
        return self.__matvec(x)

    def rmatvec(self, x):
        if self.__H is None:
            raise NotImplementedError("operator has no transpose")
        return self.__H.matvec(x)

    def to_array(self):
        n, m = self.shape
        H = np.empty((n, m), dtype=self.dtype)
        e = np.zeros(m, dtype=self.dtype)
        for j in range(m):
            e[j] = 1
            H[:, j] = self * e
            e[j] = 0
        return H

    def __call__(self, *args, **kwargs):
        return self.__mul__(*args, **kwargs)

    def dot(self, x):
        return self.__mul__(x)

    def __mul_scalar(self, a):
        if a == 0:
            return ZeroOperator(self.nargin, self.nargout)
        rm = None if self.__H is None else (lambda x: np.conj(a) * self.__H.matvec(x))
        return LinearOperator(self.nargin, self.nargout, matvec=lambda x: a * self.matvec(x),
                              rmatvec=rm, symmetric=self.symmetric,
                              dtype=np.result_type(self.dtype, type(a)))

    def __mul_linop(self, op):
        if self.nargin != op.nargout:
            raise ShapeError("Cannot multiply operators together")
        if self.__H is not None and op.T is not None:
            rm = lambda x: op.T.matvec(self.__H.matvec(x))  # noqa: E731
        else:
            rm = None
        return LinearOperator(op.nargin, self.nargout,
                              matvec=lambda x: self.matvec(op.matvec(x)), rmatvec=rm,
                              dtype=np.result_type(self.dtype, op.dtype))

    def __mul_vector(self, x):
        x = np.asanyarray(x)
        if x.ndim == 2 and x.shape[1] == 1:
            return self.matvec(x[:, 0]).reshape(-1, 1)
        if x.shape[0] != self.nargin:
            raise ShapeError("Multiplying with vector of wrong shape.")
        return self.matvec(x)

    def __mul__(self, x):
        if np.isscalar(x):
            return self.__mul_scalar(x)
        if isinstance(x, BaseLinearOperator):
            return self.__mul_linop(x)
        if isinstance(x, (np.ndarray, list, tuple)):
            return self.__mul_vector(x)
        raise ValueError("Cannot multiply")

    def __rmul__(self, x):
        if np.isscalar(x):
            return self.__mul__(x)
        raise ValueError("Cannot multiply")

    def __add__(self, other):
        if not isinstance(other, BaseLinearOperator):
            raise ValueError("Cannot add")
        if self.shape != other.shape:
            raise ShapeError("Cannot add")
        if self.__H is not None and other.T is not None:
            rm = lambda x: self.__H.matvec(x) + other.T.matvec(x)  # noqa: E731
        else:
            rm = None
        return LinearOperator(self.nargin, self.nargout,
                              matvec=lambda x: self.matvec(x) + other.matvec(x), rmatvec=rm,
                              symmetric=self.symmetric and other.symmetric,
                              dtype=np.result_type(self.dtype, other.dtype))

    def __neg__(self):
        return self * (-1)

    def __sub__(self, other):
        if not isinstance(other, BaseLinearOperator):
            raise ValueError("Cannot add")
        if self.shape != other.shape:
            raise ShapeError("Cannot add")
        if self.__H is not None and other.T is not None:
            rm = lambda x: self.__H.matvec(x) - other.T.matvec(x)  # noqa: E731
        else:
            rm = None
        return LinearOperator(self.nargin, self.nargout,
                              matvec=lambda x: self.matvec(x) - other.matvec(x), rmatvec=rm,
                              symmetric=self.symmetric and other.symmetric,
                              dtype=np.result_type(self.dtype, other.dtype))

    def __truediv__(self, a):
        if np.isscalar(a):
            return self * (1.0 / a)
        raise ValueError("Cannot divide")


class IdentityOperator(LinearOperator):
    def __init__(self, nargin, **kwargs):
        kwargs.pop("symmetric", None)
        super(IdentityOperator, self).__init__(nargin, nargin, symmetric=True,
                                               matvec=lambda x: x, **kwargs)


class DiagonalOperator(LinearOperator):
    def __init__(self, diag, **kwargs):
        diag = np.asarray(diag)
        if diag.ndim != 1:
            raise ValueError("Input must be 1-d array")
        self.__diag = diag.copy()
        kwargs.pop("symmetric", None)
        super(DiagonalOperator, self).__init__(diag.shape[0], diag.shape[0], symmetric=True,
                                               matvec=lambda x: self.__diag * x,
                                               dtype=self.__diag.dtype, **kwargs)

    @property
    def diag(self):
        return self.__diag


class ZeroOperator(LinearOperator):
    def __init__(self, nargin, nargout, **kwargs):
        kwargs.pop("matvec", None)
        kwargs.pop("rmatvec", None)
        super(ZeroOperator, self).__init__(nargin, nargout,
                                           matvec=lambda x: np.zeros(nargout),
                                           rmatvec=lambda x: np.zeros(nargin), **kwargs)
