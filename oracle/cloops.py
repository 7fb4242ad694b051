"""
ORACLE (test infrastructure, not product code): ctypes binding of oracle/weave_loops.c, the
plain-C restatement of the reference's weave.inline loops.  ``build()`` compiles it with
``gcc -O3`` (the reference's own flags, interfaces/linearoperators.py:377) into
oracle/_build/liboracle_loops.so; nothing here is used by the product path.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "weave_loops.c")
_SO = os.path.join(_HERE, "_build", "liboracle_loops.so")
_lib = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the C restatement (idempotent)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        try:
            build()
        except Exception:
            _lib = False
            return _lib
    lib = ctypes.CDLL(_SO)
    lib.orc_seq_sum.restype = ctypes.c_double
    lib.orc_dot.restype = ctypes.c_double
    _lib = lib
    return _lib


def available():
    return bool(_load())


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _pix64(pix):
    return np.ascontiguousarray(pix, dtype=np.int64)


def pointing_mult(pix, c, s, pol, v):
    pix = _pix64(pix)
    x = np.zeros(len(pix))
    _load().orc_pointing_mult(_p(pix, _i64p), _p(c, _f64p), _p(s, _f64p), ctypes.c_int64(len(pix)),
                              ctypes.c_int(pol), _p(np.ascontiguousarray(v), _f64p), _p(x, _f64p))
    return x


def pointing_rmult(pix, c, s, pol, v, npix):
    pix = _pix64(pix)
    x = np.zeros(npix * pol)
    _load().orc_pointing_rmult(_p(pix, _i64p), _p(c, _f64p), _p(s, _f64p), ctypes.c_int64(len(pix)),
                               ctypes.c_int(pol), _p(np.ascontiguousarray(v), _f64p), _p(x, _f64p))
    return x


def moments(pix, w, c, s, pol, npix):
    pix = _pix64(pix)
    out = [np.zeros(npix) for _ in range(6)]
    _load().orc_moments(_p(pix, _i64p), _p(np.ascontiguousarray(w, dtype=np.float64), _f64p),
                        _p(c, _f64p), _p(s, _f64p), ctypes.c_int64(len(pix)), ctypes.c_int(pol),
                        *[_p(o, _f64p) for o in out])
    return out  # counts, cosine, sine, cos2, sin2, sincos


def bd_apply(npix, pol, hits, c, s, c2, s2, cs, x):
    y = np.zeros(npix * pol)
    _load().orc_bd_apply(ctypes.c_int64(npix), ctypes.c_int(pol), _p(hits, _f64p), _p(c, _f64p),
                         _p(s, _f64p), _p(c2, _f64p), _p(s2, _f64p), _p(cs, _f64p),
                         _p(np.ascontiguousarray(x), _f64p), _p(y, _f64p))
    return y


def toeplitz(a, v):
    y = np.empty(len(v))
    _load().orc_toeplitz(_p(a, _f64p), ctypes.c_int64(len(a)), _p(v, _f64p),
                         ctypes.c_int64(len(v)), _p(y, _f64p))
    return y


def seq_sum(d):
    return float(_load().orc_seq_sum(_p(d, _f64p), ctypes.c_int64(len(d))))


def filter_offset(pix, d, start, end):
    pix = _pix64(pix)
    d = np.ascontiguousarray(d, dtype=np.float64)
    out = np.zeros(len(d))
    start = np.ascontiguousarray(start, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    _load().orc_filter_offset(_p(pix, _i64p), _p(d, _f64p), _p(out, _f64p), _p(start, _i64p),
                              _p(end, _i64p), ctypes.c_int64(len(start)))
    return out


def dot(a, b):
    return float(_load().orc_dot(_p(a, _f64p), _p(b, _f64p), ctypes.c_int64(len(a))))
